#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_tasks.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/kernel_bench.py --only text_qkv --stamps 2>&1 | grep -v globaltimer | tail -4
timeout 100 python scripts/kernel_bench.py --only plain_text_attn_out --stamps 2>&1 | grep -v globaltimer | tail -4
for rep in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --ops-table gpurun_out/ops_table.jsonl > gpurun_out/x$rep.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/x$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("run $rep", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], round(r["achieved"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("run $rep ERR", e)
PY
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inflight 1 2>&1 | tail -1 | cut -c1-200
