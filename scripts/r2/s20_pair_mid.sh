#!/bin/bash
# CTA-pair threshold at the batches between 64 and 512.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s20_ab.txt
for rep in 1 2; do
for b in 128 192 256 384; do
for w in 4 2; do
  VB200_PAIR_MIN_WAVES=$w timeout 300 python bench.py --batch $b --steps 60 --warmup 4 --no-cpu-baseline --dtype fp16 > $O/s20_tmp.json 2> $O/s20_tmp.err
  python - <<PY >> $O/s20_ab.txt
import json
try:
    j = json.load(open("$O/s20_tmp.json")); r = j["roofline"]
    print("rep=$rep batch=$b pair_min_waves=$w", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), "frac", round(r["frac"], 3), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("b=$b w=$w ERR", e, open("$O/s20_tmp.err").read()[-600:])
PY
done
done
done
cat $O/s20_ab.txt
