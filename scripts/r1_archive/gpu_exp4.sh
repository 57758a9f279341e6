#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("$name ERR", e)
PY
}
run medium A=1
run mediumplus VB200_PDL=mediumplus
run full VB200_PDL=full
run medium_b A=1
run mediumplus_b VB200_PDL=mediumplus
env A=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inflight 3 2>&1 | tail -1 | cut -c1-200
