#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s7b_grid.txt
for rep in 1 2; do
for pct in 100 85 75 67 60 50; do
  for fl in 2 3; do
    VB200_GRID_PCT=$pct timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $fl > $O/s7_tmp.json 2> $O/s7_tmp.err
    python - <<PY >> $O/s7b_grid.txt
import json
try:
    j = json.load(open("$O/s7_tmp.json")); r = j["roofline"]
    print("rep=$rep grid_pct=$pct inflight=$fl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("grid_pct=$pct inflight=$fl ERR", e, open("$O/s7_tmp.err").read()[-300:])
PY
  done
done
done
cat $O/s7b_grid.txt
