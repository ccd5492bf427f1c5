#!/bin/bash
# both bench arms + the central workload, no profiler (the lines that go into profiles/)
mkdir -p gpurun_out
T=${1:-fin}
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 900 python bench.py > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err
timeout 600 python bench.py --workload central4 --no-cpu-baseline > gpurun_out/${T}_central4_bench512.json 2> gpurun_out/${T}_central4_bench512.err
timeout 600 python bench.py --workload central4 --size 64 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/${T}_central4_bench64.json 2> gpurun_out/${T}_central4_bench64.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_list.log 2>&1
python - <<PY
import json
for f in ('bench_reference','bench512','central4_bench512','central4_bench64'):
    d=json.loads(open('gpurun_out/${T}_%s.json'%f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d.get('ms_per_step'), d.get('e2e') and '%.4g'%d['e2e']['value'])
PY
