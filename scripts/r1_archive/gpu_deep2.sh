#!/bin/bash
cd "$(dirname "$0")/.."
for d in 1 0; do
echo "=== VB200_DEEP=$d"
VB200_DEEP=$d timeout 100 python scripts/kernel_bench.py --only plain_ --stamps 2>&1 | grep -v globaltimer | tail -9
VB200_DEEP=$d timeout 100 python scripts/kernel_bench.py --only pool_t --stamps 2>&1 | grep -v globaltimer | tail -3
done
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "persistent" 2>&1 | tail -2
