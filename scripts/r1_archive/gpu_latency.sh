#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 1 8 64; do
for pdl in medium full; do
VB200_PDL=$pdl timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --inflight 1 --batch $b > gpurun_out/lat_${b}_$pdl.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/lat_${b}_$pdl.log").read().strip().splitlines()[-1])
    print("batch $b pdl $pdl: %.3f ms per forward, %d pairs/s; host-buffer call %.3f ms" % (j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print("batch $b pdl $pdl ERR", e)
PY
done
done
