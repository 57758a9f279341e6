#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 1 8; do
for cfg in "VB200_PDL=full" "VB200_PDL=full VB200_DEEP=1" "VB200_PDL=full VB200_DEEP=1 VB200_SPLITK=1"; do
env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --inflight 1 --batch $b > gpurun_out/lat2.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/lat2.log").read().strip().splitlines()[-1])
    print("batch $b [$cfg]: %.3f ms per forward, %d pairs/s; host-buffer call %.3f ms" % (j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print("batch $b [$cfg] ERR", e)
PY
done
done
