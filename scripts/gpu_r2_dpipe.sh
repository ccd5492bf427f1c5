#!/bin/bash
# distributed window pipeline: multi-GPU test + the bench line at N ranks (e2e pipelined vs unpipelined)
N=${1:-2}; T=${2:-r2dp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -x -k "distributed_window" > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_bench_n$N.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('N=$N value %.4g ms/step %.1f' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e.get('value', 0), 'unpipelined', (e.get('unpipelined') or {}).get('value'), e.get('pipelined_error'), d.get('parity'))
PY
tail -3 gpurun_out/${T}_bench_n$N.err
