#!/bin/bash
# multi-GPU check: tests (incl. 2-GPU bit-exactness) then weak-scaling bench at N ranks
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -2 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1])
print('N=%d value %.4g ms/step %.2f e2e %.4g launches %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
PY
