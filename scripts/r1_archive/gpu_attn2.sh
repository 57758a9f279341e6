#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_tasks.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --ops-table gpurun_out/ops_table.jsonl > gpurun_out/y$rep.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/y$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("run $rep", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], round(r["achieved"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("run $rep ERR", e)
PY
done
grep "attention\|\[64, 1024, 1024, 2\]" gpurun_out/ops_table.jsonl
