#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:'self_attention' -s 10 -c 1 \
   -o gpurun_out/prof_self_attn_img -f python scripts/kernel_bench.py --only self_attn_img --reps 20 > gpurun_out/ncu_sa.log 2>&1
echo "ncu sa exit $?"
timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:'co_attention' -s 10 -c 1 \
   -o gpurun_out/prof_co_attn -f python scripts/kernel_bench.py --only co_attn --reps 20 > gpurun_out/ncu_co.log 2>&1
echo "ncu co exit $?"
ls -la gpurun_out/*.ncu-rep
