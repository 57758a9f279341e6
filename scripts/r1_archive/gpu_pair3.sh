#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 64 512; do
 for cfg in "0 0" "2 128" "2 256"; do
  set -- $cfg
  echo "== batch $b variant $1 bn $2"
  timeout 200 python scripts/kernel_bench.py --batch $b --variant $1 --bn $2 --only img_qkv --stamps 2>&1 | tail -5
  timeout 200 python scripts/kernel_bench.py --batch $b --variant $1 --bn $2 --only text_ffn_in_gelu --stamps 2>&1 | grep -v bn256 | tail -5
 done
done
