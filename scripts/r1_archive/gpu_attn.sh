#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -2
timeout 100 python scripts/kernel_bench.py --only _attn 2>&1 | grep -v "_ln\|plain" | tail -5
