#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -12
grep "library_bar" gpurun_out/parity.jsonl | tail -1
for rep in 1 2 3; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/v$rep.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/v$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("run $rep", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("run $rep ERR", e)
PY
done
grep -v "^\[W" gpurun_out/v1.log | grep -i "vb:\|error" | head -5
