#!/bin/bash
# round 2: full GPU test suite (incl. the config-scale parity tests) + the headline bench line with its parity object
T=${1:-r2a}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=15 > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
tail -30 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_bench512.json; tail -5 gpurun_out/${T}_bench512.err
