#!/bin/bash
# Persistent grid below half of the CTA slots x batches in flight (every CTA walks 3-4 tiles; one CTA of a kernel per SM).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s14_ab.txt
for rep in 1 2; do
for cfg in "67 2" "33 2" "33 3" "33 4" "25 3" "25 4" "40 3" "50 3" "50 4" "67 3"; do
  set -- $cfg
  VB200_GRID_PCT=$1 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $2 > $O/s14_tmp.json 2> $O/s14_tmp.err
  python - <<PY >> $O/s14_ab.txt
import json
try:
    j = json.load(open("$O/s14_tmp.json")); r = j["roofline"]
    print("rep=$rep grid=$1 inflight=$2", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("grid=$1 inflight=$2 ERR", e, open("$O/s14_tmp.err").read()[-400:])
PY
done
done
cat $O/s14_ab.txt
