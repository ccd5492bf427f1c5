#!/bin/bash
N=${1:-2}; T=${2:-r2lc}
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-parity --no-secondary --steps 3 > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --scaling strong --grid 1024 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/${T}_strong_n$N.json 2> gpurun_out/${T}_strong_n$N.err
python - <<PY
import json
for f in ('n1', 'n$N', 'strong_n$N'):
    try:
        d = [json.loads(l) for l in open('gpurun_out/${T}_%s.json' % f) if l.startswith('{')][-1]
        e = d.get('e2e') or {}
        print(f, 'value %.4g' % d['value'], 'e2e %.4g' % e.get('value', 0), 'unpipelined', (e.get('unpipelined') or {}).get('value'), e.get('pipelined_error'))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
tail -2 gpurun_out/${T}_strong_n$N.err
