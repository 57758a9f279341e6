#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
echo "ops exit $?" > gpurun_out/status.txt
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "model exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 512 > gpurun_out/bench_b512.log 2>&1
timeout 600 python scripts/kernel_bench.py --stamps --only ffn_in > gpurun_out/kernel_bench.log 2>&1
tail -3 gpurun_out/pytest_ops.log; tail -3 gpurun_out/pytest_model.log | cut -c1-200
for f in plain b512; do tail -1 gpurun_out/bench_$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['families_ms'])"; done; cat gpurun_out/kernel_bench.log; cat gpurun_out/status.txt
