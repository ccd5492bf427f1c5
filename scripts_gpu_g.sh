#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/g_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_tests.log; tail -4 gpurun_out/g_tests.log
for s in 64 128; do
timeout 600 python bench.py --workload central4 --size $s --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/g_c$s.json 2> gpurun_out/g_c$s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/g_c$s.json').read().strip().splitlines()[-1])
print('central', $s, 'value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']))
PY
done
timeout 600 python bench.py --size 64 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/g_t64.json 2> gpurun_out/g_t64.err
python - <<PY
import json
d=json.loads(open('gpurun_out/g_t64.json').read().strip().splitlines()[-1])
print('teno5 64', 'value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']))
PY
