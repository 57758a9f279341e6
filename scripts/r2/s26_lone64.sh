#!/bin/bash
# Small-batch latency: 128x64 tiles (twice the CTAs, half the epilogue per CTA) under the lone-CTA configuration.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s26_ab.txt
for rep in 1 2; do
for b in 1 8; do
for bn in 128 64; do
  VB200_BN=$bn timeout 300 python bench.py --batch $b --inflight 1 --steps 300 --warmup 10 --no-cpu-baseline --dtype fp16 > $O/s26_tmp.json 2> $O/s26_tmp.err
  python - <<PY >> $O/s26_ab.txt
import json
try:
    j = json.load(open("$O/s26_tmp.json")); r = j["roofline"]
    print("rep=$rep batch=$b bn=$bn", "ms/forward", round(j["ms_per_step"], 4), "e2e ms", round(j["e2e"]["ms_per_step"], 4), r["families_ms"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("b=$b bn=$bn ERR", e, open("$O/s26_tmp.err").read()[-600:])
PY
done
done
done
cat $O/s26_ab.txt
