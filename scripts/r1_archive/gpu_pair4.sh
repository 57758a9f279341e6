#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VB200_DEBUG=1
for b in 64 512; do
 for res in 0 148; do
  echo "== batch $b pair128 resident-override $res"
  if [ $res = 0 ]; then unset VB200_PAIR_RESIDENT; else export VB200_PAIR_RESIDENT=$res; fi
  timeout 200 python scripts/kernel_bench.py --batch $b --variant 2 --bn 128 --only img_qkv --stamps 2>&1 | tail -5
  timeout 200 python scripts/kernel_bench.py --batch $b --variant 2 --bn 128 --only text_qkv --stamps 2>&1 | tail -4
 done
done
