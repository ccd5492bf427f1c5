#!/bin/bash
# repeat the 2-GPU bit-exactness test to expose ordering races in the peer-store halo exchange
mkdir -p gpurun_out
for v in fromq old; do
  if [ $v = old ]; then export OSB_NO_VISCOUS_FROM_Q=1; fi
  f=0
  for i in 1 2 3 4 5 6; do
    timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -k teno5 > gpurun_out/mgs_$v$i.log 2>&1 || { f=$((f+1)); grep -m1 "AssertionError" gpurun_out/mgs_$v$i.log | cut -c1-200; }
  done
  echo "$v failures: $f of 6"
done
