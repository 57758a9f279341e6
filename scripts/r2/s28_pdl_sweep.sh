#!/bin/bash
# Programmatic-dependent-launch scope re-measured on the final build (round 1 chose "full" with the full persistent grid).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s28_ab.txt
for rep in 1 2 3; do
for pdl in full mediumplus medium light; do
  VB200_PDL=$pdl timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 > $O/s28_tmp.json 2> $O/s28_tmp.err
  python - <<PY >> $O/s28_ab.txt
import json
try:
    j = json.load(open("$O/s28_tmp.json"))
    print("rep=$rep pdl=$pdl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("pdl=$pdl ERR", e, open("$O/s28_tmp.err").read()[-400:])
PY
done
done
cat $O/s28_ab.txt
