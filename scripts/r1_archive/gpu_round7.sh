#!/bin/bash
# 2-GPU box: 2-GPU bench, smoke, launch list of the default bench, full gpu test-suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
echo "bench2 exit $?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_ref2.log 2>&1
echo "ref2 exit $?"; tail -1 gpurun_out/bench_ref2.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
