#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; b=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batch $b > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("$name ERR", e)
PY
}
run b512_full 512 A=1
run b512_off 512 VB200_PDL=off
run b512_full2 512 A=1
run b512_off2 512 VB200_PDL=off
run b256_full 256 A=1
run b256_off 256 VB200_PDL=off
