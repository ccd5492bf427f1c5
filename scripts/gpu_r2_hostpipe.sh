#!/bin/bash
# window pipeline over a host-resident block: parity tests, then the end-to-end leg of the bench at three window sizes
T=${1:-r2hp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hostpipe.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -25 gpurun_out/${T}_tests.log
for C in 32 64 128; do
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-secondary --steps 3 --warmup 3 --e2e-chunk $C > gpurun_out/${T}_bench_c$C.json 2> gpurun_out/${T}_bench_c$C.err
  python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_bench_c$C.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('chunk $C: ms/step %.2f' % d['ms_per_step'], 'e2e %.4g' % e.get('value', 0), 'unpipelined', (e.get('unpipelined') or {}).get('value'), e.get('pipelined_error'), e.get('call'))
PY
  tail -3 gpurun_out/${T}_bench_c$C.err
done
