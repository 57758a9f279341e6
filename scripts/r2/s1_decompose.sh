#!/bin/bash
# Round 2, GPU session 1: (a) what bounds each GEMM -- the same kernel with no MMA (1) / no stores (2) / no operand loads (4);
# (b) bench lines of the current build in both 16-bit formats; (c) compute-sanitizer over op tests and a tiny forward.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_env.txt
K=text_qkv,text_ffn_in_gelu,plain_text_ffn_out,img_qkv,plain_img_out,plain_text_attn_out
for b in 64 512; do
  echo "== batch $b" >> $O/s1_decompose.txt
  timeout 300 python scripts/kernel_bench.py --batch $b --debug 0,1,2,4,5,3 --only $K --stamps --reps 100 >> $O/s1_decompose.txt 2>&1
done
for dt in bf16 fp16; do
  timeout 300 python bench.py --steps 100 --warmup 5 --dtype $dt --no-cpu-baseline > $O/s1_bench_$dt.json 2> $O/s1_bench_$dt.err
done
# sanitizer: a slice of the op tests (small shapes) and one tiny-model forward
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="(test_linear_bias and 200-384) or (test_linear_bias and 256-256) or (test_linear_residual_layernorm and 300-256) or (test_self_attention and 3-31) or (test_co_attention and 3-31) or (test_layernorm_row_kernel and 37-128)"
for tool in memcheck racecheck synccheck; do
  timeout 300 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$SEL" > $O/s1_san_${tool}_ops.txt 2>&1
  echo "exit $?" >> $O/s1_san_${tool}_ops.txt
  timeout 300 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "test_tiny_model_all_outputs and 2-30-36" > $O/s1_san_${tool}_tiny.txt 2>&1
  echo "exit $?" >> $O/s1_san_${tool}_tiny.txt
done
tail -4 $O/s1_san_*.txt
tail -3 $O/s1_decompose.txt; cut -c1-300 $O/s1_bench_bf16.json
