#!/bin/bash
# quick perf probe: bench 256^3 + ncu on the flux kernels
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench256.json 2> gpurun_out/${TAG}_bench256.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 6 -c 3 -o gpurun_out/${TAG}_flux python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench256.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.2f'%(d['value'],d['ms_per_step']), d['roofline']['families_ms'], 'finite', d['config']['finite'])
PY
