#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:'gemm_persistent' -s 30 -c 1 \
   -o gpurun_out/prof_gemm_ffn_out -f python scripts/kernel_bench.py --only plain_text_ffn_out --reps 20 > gpurun_out/ncu_ffn_out.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_gemm_ffn_out.ncu-rep
