#!/bin/bash
# GPU run 3: v2 staged flux kernels: tests, bench, ncu
mkdir -p gpurun_out
TAG=${1:-r3}
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench256.json 2> gpurun_out/${TAG}_bench256.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench512.json 2> gpurun_out/${TAG}_bench512.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_viscous3d" -s 8 -c 4 -o gpurun_out/${TAG}_flux python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log; python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench256.json','gpurun_out/${TAG}_bench512.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g ms/step %.2f'%(d['value'],d['ms_per_step']), 'roofline frac', d['roofline'] and round(d['roofline']['frac'],3), d['roofline'] and d['roofline']['families_ms'], 'e2e', d.get('e2e') and '%.4g'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
