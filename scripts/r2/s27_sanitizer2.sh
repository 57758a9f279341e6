#!/bin/bash
# compute-sanitizer over what round 2 added after GPU session 1: chained FFN launch, three CTAs per SM (MODE 6), one CTA per SM with a
# 6-stage ring (MODE 7, default for small forwards, also 64-wide), the fold / split GEMM modes, the fp32 attention, region pack, and a
# tiny-model forward (which now runs MODE 7) with and without the timeline hook.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="(test_gemm_chain and 300-256-384) or (test_linear_three_ctas_per_sm and 200-384-768) or (test_linear_three_ctas_per_sm and 256-256-1024) or (test_linear_lone_cta_64_wide and 31-768-768) or (test_linear_lone_cta_64_wide and 288-1024-1024) or (test_linear_split_operands and 300-384-768) or (test_linear_split_operands and 77-64-128) or (test_layernorm_fold_chain and 300-256-384) or (test_attention_f32 and 3-31-36)"
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$SEL" > $O/s27_san_${tool}_ops.txt 2>&1
  echo "exit $?" >> $O/s27_san_${tool}_ops.txt
  timeout 600 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "timeline or (cached_states and tiny_fp16) or (fp32x_tiny and 2-30-36)" > $O/s27_san_${tool}_tiny.txt 2>&1
  echo "exit $?" >> $O/s27_san_${tool}_tiny.txt
done
for f in $O/s27_san_*.txt; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|exit" $f | tail -6; done
