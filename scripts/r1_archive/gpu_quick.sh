#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py --stamps > gpurun_out/kb_stamps.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1
cat gpurun_out/kb_stamps.log; tail -1 gpurun_out/bench_plain.log | cut -c1-200
