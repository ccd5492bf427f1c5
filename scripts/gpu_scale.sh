#!/bin/bash
# weak-scaling bench at N ranks exactly as the driver launches it (headline workload), plus the central-4 workload
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err
cat gpurun_out/scale_n${N}.json; tail -2 gpurun_out/scale_n${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload central4 --steps 10 --warmup 3 --no-e2e > gpurun_out/scale_central4_n${N}.json 2> gpurun_out/scale_central4_n${N}.err
cat gpurun_out/scale_central4_n${N}.json
