#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 8 4 2 1; do
VB200_LN_ROWS=$r timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ln$r.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/ln$r.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("rows $r", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"])
except Exception as e:
    print("rows $r ERR", e)
PY
done
