#!/bin/bash
# flux sweeps: tests with the product library, then bench lines of the product library and of experimental builds
# (opensbli_b200/libosbli_b200_<tag>.so, OSB_B200_LIB), then one full ncu capture of the product's sweep kernels
T=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
for lib in opensbli_b200/libosbli_b200_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libosbli_b200_//')
  OSB_B200_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/${T}_bench512_${tag}.json 2> gpurun_out/${T}_bench512_${tag}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${T}_bench512*.json')):
    try:
        d = json.load(open(f))
        print(f, 'ms/step %.2f' % d['ms_per_step'], 'flux launch ms %.3f' % d['roofline']['launch_ms'], d['roofline']['families_ms'], (d.get('parity') or {}).get('max_rel_err'))
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flux3" -s 6 -c 3 -o gpurun_out/${T}_flux3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
