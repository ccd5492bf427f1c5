#!/bin/bash
# generic path: GPU tests, then hand-written vs generic speed of the bench workloads at 128^3 and 256^3
T=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_app_run.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
for N in 128 256; do
  timeout 900 python scripts/generic_speed.py $N > gpurun_out/${T}_generic_speed_$N.json 2> gpurun_out/${T}_generic_speed_$N.err
  tail -c 1500 gpurun_out/${T}_generic_speed_$N.json; tail -3 gpurun_out/${T}_generic_speed_$N.err
done
