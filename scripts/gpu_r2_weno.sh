#!/bin/bash
# WENO5 weights over a common denominator: parity tests with the product library, then ms/step of both forms
T=${1:-r2w5}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 600 python scripts/weno_speed.py > gpurun_out/${T}_weno_speed_product.json 2> gpurun_out/${T}_weno_speed_product.err; cat gpurun_out/${T}_weno_speed_product.json; tail -2 gpurun_out/${T}_weno_speed_product.err
for lib in opensbli_b200/libosbli_b200_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libosbli_b200_//')
  OSB_B200_LIB=$PWD/$lib timeout 600 python scripts/weno_speed.py > gpurun_out/${T}_weno_speed_$tag.json 2> gpurun_out/${T}_weno_speed_$tag.err; cat gpurun_out/${T}_weno_speed_$tag.json; tail -2 gpurun_out/${T}_weno_speed_$tag.err
done
