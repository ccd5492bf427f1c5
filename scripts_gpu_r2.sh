#!/bin/bash
# GPU run 2 (2 GPUs): all gpu tests incl. multi-GPU + app runs, then the weak-scaling bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size 256 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_256.json 2> gpurun_out/r2_bench_n2_256.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_512.json 2> gpurun_out/r2_bench_n2_512.err
tail -8 gpurun_out/r2_tests.log; cat gpurun_out/r2_bench_n2_256.json; tail -3 gpurun_out/r2_bench_n2_256.err; cat gpurun_out/r2_bench_n2_512.json; tail -3 gpurun_out/r2_bench_n2_512.err
