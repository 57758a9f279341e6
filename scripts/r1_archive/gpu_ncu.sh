#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in text_attn_out_ln text_ffn_in_gelu; do
  timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:'gemm_persistent' -s 30 -c 1 \
     -o gpurun_out/prof_$k -f python scripts/kernel_bench.py --only $k --reps 20 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?" >> gpurun_out/status.txt
done
ls -la gpurun_out/*.ncu-rep
