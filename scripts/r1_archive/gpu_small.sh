#!/bin/bash
cd "$(dirname "$0")/.."
timeout 100 python scripts/kernel_bench.py --only pool_t --stamps 2>&1 | tail -4
timeout 100 python scripts/kernel_bench.py --only vqa_fc3 --stamps 2>&1 | tail -4
