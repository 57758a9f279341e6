#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 1200 python -m pytest tests/test_gpu_tasks.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_tasks.log 2>&1
echo "tasks exit $?" > gpurun_out/status.txt
timeout 900 python scripts/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
echo "sweep exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1
tail -12 gpurun_out/pytest_tasks.log | cut -c1-200; cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err; tail -1 gpurun_out/bench_plain.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline'])"; cat gpurun_out/status.txt
