#!/bin/bash
# N GPUs of one box: bit-exactness tests of the slab decomposition (incl. uneven slabs, the app-level runner under torchrun),
# then weak- and strong-scaling bench lines exactly as the driver launches them
N=${1:-4}; T=${2:-r2m$N}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $2 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_$1.json") if l.startswith("{")][-1])
    print('$1', 'grid', d['config']['grid'], 'ms/step %.2f' % d['ms_per_step'], 'value %.4g' % d['value'], 'sync ms', round(d['roofline']['families_ms'].get('sync', 0), 2), 'e2e', (d.get('e2e') or {}).get('value'), 'parity', d.get('parity'))
except Exception as e:
    print('$1', 'failed', e)
PY
}
run weak "--no-e2e"
run strong1024 "--scaling strong --grid 1024 --no-e2e"
run strong512 "--scaling strong --grid 512 --no-e2e --no-parity"
