#!/bin/bash
# Lone-CTA GEMMs with TWO batches in flight (throughput mode at small batch): does one CTA per SM cost concurrency?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s25_ab.txt
for rep in 1 2; do
for b in 4 8 16; do
for lone in 0 320 640; do
  VB200_LONE_ROWS=$lone timeout 300 python bench.py --batch $b --steps 300 --warmup 10 --no-cpu-baseline --dtype fp16 > $O/s25_tmp.json 2> $O/s25_tmp.err
  python - <<PY >> $O/s25_ab.txt
import json
try:
    j = json.load(open("$O/s25_tmp.json"))
    print("rep=$rep batch=$b inflight=2 lone_rows=$lone", "ms/step", round(j["ms_per_step"], 4), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]))
except Exception as e:
    print("b=$b lone=$lone ERR", e, open("$O/s25_tmp.err").read()[-600:])
PY
done
done
done
cat $O/s25_ab.txt
