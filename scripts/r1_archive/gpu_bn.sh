#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for bn in 128 256; do
  for b in 64 512; do
    VB200_BN=$bn timeout 300 python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline > gpurun_out/bn${bn}_b${b}.log 2>&1
    python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bn${bn}_b${b}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("bn$bn b$b", round(j["value"]), round(j["ms_per_step"],3), "gemm TF", round(r["achieved"]), r["families_ms"], r["largest_gemm"])
except Exception as e:
    print("bn$bn b$b ERR", e)
PY
  done
done
