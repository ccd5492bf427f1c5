#!/bin/bash
TAG=${1:-t}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log
tail -40 gpurun_out/${TAG}_tests.log
