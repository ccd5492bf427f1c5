#!/bin/bash
# the headline workload from a developed state (precursor run on 256^3 tiled 2x2x2): ms/step and share of waves on the slow path
T=${1:-r2dev}
mkdir -p gpurun_out
for TT in 4 8; do
  timeout 900 python bench.py --state developed --develop-time $TT --no-cpu-baseline --no-e2e --no-secondary --no-parity --steps 5 > gpurun_out/${T}_t$TT.json 2> gpurun_out/${T}_t$TT.err
  python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_t$TT.json') if l.startswith('{')][-1]
print('t=$TT', 'ms/step %.2f' % d['ms_per_step'], 'value %.4g' % d['value'], d['state'])
PY
  tail -2 gpurun_out/${T}_t$TT.err
done
