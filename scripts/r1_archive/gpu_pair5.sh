#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pr in 0 128; do
  for b in 64 128 512; do
    VB200_PAIR=$pr timeout 300 python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline > gpurun_out/pair${pr}_b${b}.log 2>&1
    python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/pair${pr}_b${b}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("pair$pr b$b", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"], r["largest_gemm"]["tflops"])
except Exception as e:
    print("pair$pr b$b ERR", e)
PY
  done
done
VB200_PAIR=128 timeout 300 python bench.py --steps 30 --warmup 5 --inflight 1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
