#!/bin/bash
T=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hostpipe.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
for C in ramp 64; do
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-secondary --steps 3 --warmup 3 --e2e-chunk $C > gpurun_out/${T}_c$C.json 2> gpurun_out/${T}_c$C.err
  python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_c$C.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('chunk $C: e2e %.4g' % e.get('value', 0), e.get('pipelined_error'), e.get('call', '')[-90:])
PY
done
