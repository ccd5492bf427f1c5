#!/bin/bash
# central-4 workload (BASELINE config 2): parity tests touching the central path, bench at three sizes, ncu of the stage kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "central or sym or tgv or matrix or graph or e2e" > gpurun_out/c_tests.log 2>&1; tail -3 gpurun_out/c_tests.log
for s in 64 256 512; do
timeout 600 python bench.py --workload central4 --size $s --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c_bench$s.json 2> gpurun_out/c_bench$s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c_bench$s.json').read().strip().splitlines()[-1])
print($s, 'value %.4g ms/step %.3f hbm frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']), {k: round(v,2) for k,v in d['roofline']['families_ms'].items()})
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_central -s 2 -c 1 -o gpurun_out/c_central python bench.py --workload central4 --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu.log 2>&1
