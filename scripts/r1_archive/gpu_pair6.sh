#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "pair or persistent" 2>&1 | tail -3
for b in 64 512; do
  echo "== batch $b pair128"
  timeout 200 python scripts/kernel_bench.py --batch $b --variant 2 --bn 128 --only img_qkv --stamps 2>&1 | grep -v globaltimer | tail -3
  timeout 200 python scripts/kernel_bench.py --batch $b --variant 2 --bn 128 --only text_ffn_in_gelu --stamps 2>&1 | grep -v "bn256\|globaltimer" | tail -3
done
for pr in auto 128; do
  for b in 64 512; do
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
