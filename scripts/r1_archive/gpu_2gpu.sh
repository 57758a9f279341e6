#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
echo "bench2 exit $?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-260
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref2.log 2>&1
echo "ref2 exit $?"; tail -1 gpurun_out/bench_ref2.log | cut -c1-200
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
