#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s9_balance.txt
for rep in 1 2; do
for cfg in "67 1" "67 0" "75 1" "100 1" "50 1"; do
  set -- $cfg
  for fl in 2; do
    VB200_GRID_PCT=$1 VB200_GRID_BALANCE=$2 timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $fl > $O/s9_tmp.json 2> $O/s9_tmp.err
    python - <<PY >> $O/s9_balance.txt
import json
try:
    j = json.load(open("$O/s9_tmp.json")); r = j["roofline"]
    print("rep=$rep pct=$1 balance=$2 inflight=$fl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), "full", round(r["achieved_full_grid"]), j["clocks"]["sm_mhz"])
except Exception as e:
    print("pct=$1 balance=$2 ERR", e, open("$O/s9_tmp.err").read()[-300:])
PY
  done
done
done
cat $O/s9_balance.txt
