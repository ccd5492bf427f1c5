#!/bin/bash
# multi-GPU evidence: the multi-GPU tests, then the bench line at N ranks (weak scaling, pipelined e2e)
N=${1:-4}; T=${2:-r2mf}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_bench_n$N.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('N=$N value %.4g ms/step %.1f' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e.get('value', 0), 'unpipelined', (e.get('unpipelined') or {}).get('value'), e.get('pipelined_error'), (d.get('parity') or {}).get('multi_gpu_bit_identical'), (d.get('secondary') or {}).get('value'))
PY
tail -2 gpurun_out/${T}_bench_n$N.err
