#!/bin/bash
# Round 2, GPU session 2: new op tests, the round-2 model/API tests, the conv MMA-issue experiment, the default bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "split or attention_f32 or outliers or layernorm" > $O/s2_ops.txt 2>&1; echo "exit $?" >> $O/s2_ops.txt
timeout 1500 python -m pytest tests/test_gpu_round2.py -q -m gpu > $O/s2_round2.txt 2>&1; echo "exit $?" >> $O/s2_round2.txt
K=text_qkv,plain_text_ffn_out,img_qkv,plain_text_attn_out
timeout 300 python scripts/kernel_bench.py --batch 64 --debug 0,8,4,12,5,13 --only $K --stamps --reps 100 > $O/s2_conv.txt 2>&1
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/s2_bench.json 2> $O/s2_bench.err
timeout 900 python -m pytest tests/test_gpu_tasks.py -x -q -m gpu > $O/s2_tasks.txt 2>&1; echo "exit $?" >> $O/s2_tasks.txt
tail -n 15 $O/s2_ops.txt; tail -n 40 $O/s2_round2.txt; tail -n 5 $O/s2_tasks.txt; cut -c1-400 $O/s2_bench.json; tail -3 $O/s2_bench.err
