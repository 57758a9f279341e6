#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1
echo "kb exit $?" > gpurun_out/status.txt
# full ncu capture of three representative kernels (one launch each, after warm-up)
for k in text_ffn_out_ln text_ffn_in_gelu self_attn_text; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attention' -s 6 -c 1 \
     -o gpurun_out/prof_$k -f python scripts/kernel_bench.py --only $k --reps 3 --sets 1 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?" >> gpurun_out/status.txt
done
cat gpurun_out/kernel_bench.log; cat gpurun_out/status.txt
