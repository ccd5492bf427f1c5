#!/bin/bash
# round-2 evidence run (1 GPU): tests, smoke, both bench arms (smooth + perturbed state), launch list, full ncu captures of the
# dominant kernels (flux sweeps, viscous stage kernel) and of the Central-4 stage kernel
T=${1:-r2fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=8 > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
OSB_PROFILE_LIST=1 timeout 900 python bench.py > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
timeout 600 python bench.py --state perturbed --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench512_perturbed.json 2> gpurun_out/${T}_bench512_perturbed.err
timeout 600 python bench.py --workload central4 --size 64 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/${T}_central4_bench64.json 2> gpurun_out/${T}_central4_bench64.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/${T}_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_flux3|k_viscous3d" -s 8 -c 4 -o gpurun_out/${T}_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_central3d -s 4 -c 1 -o gpurun_out/${T}_central python bench.py --workload central4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_central.log 2>&1
tail -12 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench_reference.json | cut -c1-400; grep '^{' gpurun_out/${T}_bench512.json | cut -c1-300
