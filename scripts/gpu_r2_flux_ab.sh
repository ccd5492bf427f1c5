#!/bin/bash
# A/B of the flux sweeps: second generation (OSB_FLUX_V2=1) vs third generation (default), same box; tests first
T=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
OSB_FLUX_V2=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_bench512_v2.json 2> gpurun_out/${T}_bench512_v2.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench512_v3.json 2> gpurun_out/${T}_bench512_v3.err
python - <<'PY'
import json
for v in ('v2','v3'):
    try:
        d = json.load(open('gpurun_out/%s_bench512_%s.json' % ("'"$T"'".strip("'"), v)))
        print(v, 'ms/step %.2f' % d['ms_per_step'], 'flux launch ms %.3f' % d['roofline']['launch_ms'], d['roofline']['families_ms'], d.get('parity'))
    except Exception as e:
        print(v, 'failed', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flux3" -s 6 -c 3 -o gpurun_out/${T}_flux3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
