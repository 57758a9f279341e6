#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("$name ERR", e)
PY
}
run early1 A=1
run early0 VB200_EARLYW=0
run early1b A=1
run early0b VB200_EARLYW=0
