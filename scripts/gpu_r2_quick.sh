#!/bin/bash
# tests + headline bench line (with per-launch times)
T=${1:-r2q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
OSB_PROFILE_LIST=1 timeout 600 python bench.py --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench512.json'))
print('ms/step %.2f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e', (d.get('e2e') or {}).get('value'), {k: round(v, 2) for k, v in d['roofline']['families_ms'].items() if v}, d.get('parity'))
PY
grep -h "osb_profile" gpurun_out/${T}_bench512.err | head -5
