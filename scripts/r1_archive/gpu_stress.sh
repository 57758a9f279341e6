#!/bin/bash
# longer runs of the default configuration: any watchdog hit / hang / non-finite output shows up here
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2 3; do
timeout 400 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline > gpurun_out/stress$rep.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/stress$rep.log").read().strip().splitlines()[-1])
    print("stress $rep", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), j["clocks"])
except Exception as e:
    print("stress $rep ERR", e)
PY
grep -c "vb:" gpurun_out/stress$rep.log
done
timeout 400 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline --inflight 3 2>&1 | tail -1 | cut -c1-200
