#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -12
grep "library_bar" gpurun_out/parity.jsonl | tail -1
for cfg in "1 0" "0 0" "1 2" "0 2"; do
set -- $cfg
VB200_PREFETCH=$1 VB200_DEEP=$2 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$1_$2.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/ab_$1_$2.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("prefetch $1 deep $2", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("prefetch $1 deep $2 ERR", e)
PY
done
