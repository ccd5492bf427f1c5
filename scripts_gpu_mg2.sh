#!/bin/bash
N=2
mkdir -p gpurun_out
for mode in fused nofused fused nofused; do
  if [ $mode = nofused ]; then export OSB_NO_FUSED_PUSH=1; else unset OSB_NO_FUSED_PUSH; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/mgx.json 2> gpurun_out/mgx.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/mgx.json').read().strip().splitlines()[-1])
print('$mode', 'ms/step %.2f'%d['ms_per_step'], {k: round(v,1) for k,v in d['families_ms_rank0'].items()})
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/mgx1.json 2> gpurun_out/mgx1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/mgx1.json').read().strip().splitlines()[-1])
print('single', 'ms/step %.2f'%d['ms_per_step'], {k: round(v,1) for k,v in d['families_ms_rank0'].items()})
PY
