#!/bin/bash
# Step A/B: persistent-grid size (VB200_GRID_PCT) x tile width (VB200_BN192) x batches in flight.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s7_grid.txt
for cfg in "100 0" "50 0" "75 0" "50 1" "75 1" "100 1"; do
  set -- $cfg
  for fl in 2 1 3; do
    VB200_GRID_PCT=$1 VB200_BN192=$2 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $fl > $O/s7_tmp.json 2> $O/s7_tmp.err
    python - <<PY >> $O/s7_grid.txt
import json
try:
    j = json.load(open("$O/s7_tmp.json")); r = j["roofline"]
    print("grid_pct=$1 bn192=$2 inflight=$fl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"], j["clocks"]["sm_mhz"])
except Exception as e:
    print("grid_pct=$1 bn192=$2 inflight=$fl ERR", e, open("$O/s7_tmp.err").read()[-300:])
PY
  done
done
cat $O/s7_grid.txt
