#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -2
timeout 100 python scripts/kernel_bench.py --only text_qkv --stamps 2>&1 | grep -v globaltimer | tail -3
timeout 100 python scripts/kernel_bench.py --only img_qkv --stamps 2>&1 | grep -v globaltimer | tail -3
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --ops-table gpurun_out/ops_table.jsonl > gpurun_out/r9.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --inflight 1 > gpurun_out/r9_if1.log 2>&1
python - <<'PY'
import json
for n in ("r9", "r9_if1"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
grep gemm gpurun_out/ops_table.jsonl | head -12
