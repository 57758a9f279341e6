#!/bin/bash
# One GPU session: environment, parity tests, smoke, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,power.limit --format=csv
  nproc; lscpu | grep -E "Model name|^CPU\(s\)"
} > gpurun_out/env.txt 2>&1
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -rA --timeout=240 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
echo "ops exit $?" >> gpurun_out/status.txt
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -rA --timeout=600 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "model exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/status.txt
tail -5 gpurun_out/pytest_ops.log; tail -5 gpurun_out/pytest_model.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log; cat gpurun_out/status.txt
