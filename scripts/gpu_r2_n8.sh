#!/bin/bash
# 8-GPU weak-scaling bench line with the final code (1024^3 in 8 slabs), pipelined e2e leg included
T=${1:-r2n8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_bench_n8.json') if l.startswith('{')][-1]
e = d.get('e2e') or {}
print('N=8 value %.4g ms/step %.1f' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e.get('value', 0), 'unpipelined', (e.get('unpipelined') or {}).get('value'), e.get('pipelined_error'), (d.get('parity') or {}).get('multi_gpu_bit_identical'), (d.get('secondary') or {}).get('value'))
PY
tail -2 gpurun_out/${T}_bench_n8.err
