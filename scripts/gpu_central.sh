#!/bin/bash
# central-4 workload (BASELINE config 2): bench at 512^3 and 64^3 + full ncu capture of the stage kernel at 512^3
mkdir -p gpurun_out
timeout 600 python bench.py --workload central4 --no-cpu-baseline > gpurun_out/c2_central4_bench512.json 2> gpurun_out/c2_central4_bench512.err
timeout 600 python bench.py --workload central4 --size 64 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/c2_central4_bench64.json 2> gpurun_out/c2_central4_bench64.err
for s in 512 64; do python - <<PY
import json
d=json.loads(open('gpurun_out/c2_central4_bench$s.json').read().strip().splitlines()[-1])
print($s, 'value %.4g ms/step %.3f frac %.3f launch_ms %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['launch_ms']))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_central3d -s 4 -c 1 -o gpurun_out/c2_central python bench.py --workload central4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2_ncu_central.log 2>&1
