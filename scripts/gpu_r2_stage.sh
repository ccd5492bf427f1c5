#!/bin/bash
# stage kernels with the TMA plane pipeline: tests, A/B against the register-staged kernels (OSB_NO_STAGE_TMA=1), ncu
T=${1:-r2s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
OSB_PROFILE_LIST=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
OSB_PROFILE_LIST=1 OSB_NO_STAGE_TMA=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 5 > gpurun_out/${T}_bench512_notma.json 2> gpurun_out/${T}_bench512_notma.err
python - <<PY
import json
for v in ('', '_notma'):
    try:
        d = json.loads([l for l in open('gpurun_out/${T}_bench512%s.json' % v) if l.startswith('{')][-1])
        print(v or 'tma', 'ms/step %.2f' % d['ms_per_step'], {k: round(x, 2) for k, x in d['roofline']['families_ms'].items() if x}, 'secondary', (d.get('secondary') or {}).get('ms_per_step'), (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e:
        print(v, 'failed', e)
PY
grep -h "osb_profile" gpurun_out/${T}_bench512.err | head -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_viscous3d|k_central3d" -s 2 -c 2 -o gpurun_out/${T}_stage python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
