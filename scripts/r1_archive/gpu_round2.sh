#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/status.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
echo "ops exit $?" >> gpurun_out/status.txt
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "model exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1
VB200_FUSED_LN=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fusedln.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --pdl > gpurun_out/bench_pdl.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 512 > gpurun_out/bench_b512.log 2>&1
timeout 600 python scripts/kernel_bench.py --only ffn_in > gpurun_out/kernel_bench.log 2>&1
tail -4 gpurun_out/pytest_ops.log; tail -8 gpurun_out/pytest_model.log | cut -c1-200
for f in plain fusedln pdl b512; do tail -1 gpurun_out/bench_$f.log | cut -c1-200; done; cat gpurun_out/kernel_bench.log; cat gpurun_out/status.txt
