#!/bin/bash
cd "$(dirname "$0")/.."
for bn in 64 128 256; do
echo "=== bn $bn"
VB200_DEEP=0 timeout 100 python scripts/kernel_bench.py --only pool_t --bn $bn --stamps 2>&1 | grep -v globaltimer | tail -3
done
