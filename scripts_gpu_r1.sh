#!/bin/bash
# GPU run 1: parity tests, smoke, bench (256^3 then 512^3), ncu launch list and one full capture of the flux kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt 2>&1
nproc >> gpurun_out/r1_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r1_gpu.txt; free -g | head -2 >> gpurun_out/r1_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r1_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --size 256 --steps 5 --warmup 3 > gpurun_out/r1_bench256.json 2> gpurun_out/r1_bench256.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench512.json 2> gpurun_out/r1_bench512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 6 -c 3 -o gpurun_out/r1_flux python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_tests.log; cat gpurun_out/r1_smoke.log | tail -3; cat gpurun_out/r1_bench256.json; cat gpurun_out/r1_bench512.json
