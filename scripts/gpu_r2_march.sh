#!/bin/bash
# marching TMA y/z sweeps: tests, A/B against the tile kernels (OSB_NO_FLUX_MARCH=1) with per-launch times, ncu of the march kernels
T=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
OSB_PROFILE_LIST=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
OSB_PROFILE_LIST=1 OSB_NO_FLUX_MARCH=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/${T}_bench512_tile.json 2> gpurun_out/${T}_bench512_tile.err
for lib in opensbli_b200/libosbli_b200_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libosbli_b200_//')
  OSB_PROFILE_LIST=1 OSB_B200_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/${T}_bench512_${tag}.json 2> gpurun_out/${T}_bench512_${tag}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${T}_bench512*.json')):
    try:
        d = json.load(open(f))
        print(f, 'ms/step %.2f' % d['ms_per_step'], 'flux launch ms %.3f' % d['roofline']['launch_ms'], {k: round(v, 2) for k, v in d['roofline']['families_ms'].items() if v}, (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e:
        print(f, 'failed', e)
PY
grep -h "osb_profile" gpurun_out/${T}_bench512.err | head -8
grep -h "osb_profile" gpurun_out/${T}_bench512_tile.err | head -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flux3" -s 6 -c 3 -o gpurun_out/${T}_flux3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
