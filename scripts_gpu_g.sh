#!/bin/bash
# parity tests + quick headline bench (512^3 TENO5, no CPU baseline / e2e legs) + A/B against the separate-prim path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/g_tests.log 2>&1; tail -2 gpurun_out/g_tests.log
for v in new old; do
  if [ $v = old ]; then export OSB_NO_VISCOUS_FROM_Q=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/g_bench_$v.json 2> gpurun_out/g_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/g_bench_$v.json').read().strip().splitlines()[-1])
print('$v', 'value %.4g ms/step %.2f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k: round(x,2) for k,x in d['roofline']['families_ms'].items()})
PY
done
