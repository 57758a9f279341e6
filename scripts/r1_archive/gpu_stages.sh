#!/bin/bash
# DEEP variant (two MMA warps) with the ring cut to 3 / 4 stages: is a lone CTA still faster than 536 cycles per k-block?
cd "$(dirname "$0")/.."
for st in 3 4 6; do
  sed -i "s/static constexpr int kStages = kMinBlocks == 2 ? 3 : .*/static constexpr int kStages = kMinBlocks == 2 ? 3 : (DEEP ? $st : (kFit > 8 ? 8 : kFit));/" vilbert-multi-task_b200/csrc/gemm_persistent.cuh
  make -C vilbert-multi-task_b200/csrc -j8 > /dev/null 2>&1 || { echo "build failed"; exit 1; }
  echo "=== DEEP stages $st"
  VB200_DEEP=1 timeout 100 python scripts/kernel_bench.py --only plain_text_ffn_out --stamps 2>&1 | grep -v globaltimer | tail -3
  VB200_DEEP=1 timeout 100 python scripts/kernel_bench.py --only plain_img_out --stamps 2>&1 | grep -v globaltimer | tail -3
done
