#!/bin/bash
# parity tests + quick benches: headline 512^3 TENO5 and central-4 512^3 (no CPU baseline / e2e legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/g_tests.log 2>&1; tail -2 gpurun_out/g_tests.log
for w in teno5 central4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/g_bench_$w.json 2> gpurun_out/g_bench_$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/g_bench_$w.json').read().strip().splitlines()[-1])
print('$w', 'value %.4g ms/step %.2f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k: round(x,2) for k,x in d['roofline']['families_ms'].items()})
PY
done
