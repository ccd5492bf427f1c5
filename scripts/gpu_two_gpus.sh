#!/bin/bash
# 2-GPU: bit-exactness tests (run twice) + weak-scaling bench of both workloads
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -1; done
for w in teno5 central4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 --no-e2e > gpurun_out/mg3_$w.json 2> gpurun_out/mg3_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/mg3_$w.json').read().strip().splitlines()[-1])
print('$w N=%d value %.4g ms/step %.2f launches %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['gpu_launches']), {k: round(v,1) for k,v in d.get('families_ms_rank0',{}).items()})
PY
done
