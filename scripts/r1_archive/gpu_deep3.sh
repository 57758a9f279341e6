#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "persistent" 2>&1 | tail -3
timeout 100 python scripts/kernel_bench.py --only plain_ --stamps 2>&1 | grep -v globaltimer | tail -9
timeout 100 python scripts/kernel_bench.py --only pool_t --stamps 2>&1 | grep -v globaltimer | tail -3
timeout 100 python scripts/kernel_bench.py --only vqa_fc3 --stamps 2>&1 | tail -4
