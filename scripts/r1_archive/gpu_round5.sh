#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "model exit $?" > gpurun_out/status.txt
for n in 1 2 3 4; do timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --inflight $n > gpurun_out/bench_if$n.log 2>&1; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 512 --inflight 2 > gpurun_out/bench_b512.log 2>&1
tail -3 gpurun_out/pytest_model.log | cut -c1-200
for f in if1 if2 if3 if4 b512; do tail -1 gpurun_out/bench_$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['families_ms'])" || tail -3 gpurun_out/bench_$f.log; done; cat gpurun_out/status.txt
