#!/bin/bash
# pipelined end-to-end leg: window size x number of contexts
T=${1:-r2e}
mkdir -p gpurun_out
for CFG in "48 3" "64 4" "96 3" "32 4" "64 2"; do
  set -- $CFG
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-secondary --steps 3 --warmup 3 --e2e-chunk $1 --e2e-contexts $2 > gpurun_out/${T}_c$1_x$2.json 2> gpurun_out/${T}_c$1_x$2.err
  python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_c$1_x$2.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('chunk $1 contexts $2: e2e %.4g' % e.get('value', 0), e.get('pipelined_error'))
PY
done
