#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py --stamps --only text_qkv > gpurun_out/kb_stamps.log 2>&1
timeout 600 python scripts/kernel_bench.py --stamps --only text_attn_out_ln >> gpurun_out/kb_stamps.log 2>&1
timeout 600 python scripts/kernel_bench.py --stamps --only pool_t >> gpurun_out/kb_stamps.log 2>&1
cat gpurun_out/kb_stamps.log
