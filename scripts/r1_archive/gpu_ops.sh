#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ops-table gpurun_out/ops_table_b64.jsonl > gpurun_out/ops_bench.log 2>&1
cat gpurun_out/ops_table_b64.jsonl
timeout 100 python scripts/kernel_bench.py --only attn 2>&1 | tail -4
