#!/bin/bash
# model tests + default bench (1 GPU) after the LN template / pipelined e2e change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -x -q -m gpu -k "model or layernorm" 2>&1 | tail -5
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r6.log 2> gpurun_out/bench_r6.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r6.log
timeout 300 python bench.py --steps 40 --warmup 5 --inflight 1 --no-cpu-baseline > gpurun_out/bench_r6_if1.log 2>&1; python - <<'PY'
import json
for f in ("bench_r6", "bench_r6_if1"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, round(j["value"]), j["ms_per_step"], "e2e", round(j["e2e"]["value"]), j["roofline"].get("family_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
