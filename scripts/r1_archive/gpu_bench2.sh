#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1
echo "bench1 exit $?" > gpurun_out/status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
echo "bench2 exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "ref exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/status.txt
tail -1 gpurun_out/bench_1gpu.log; tail -2 gpurun_out/bench_2gpu.log | cut -c1-600; tail -1 gpurun_out/bench_ref.log | cut -c1-400; tail -2 gpurun_out/smoke.log; cat gpurun_out/status.txt
