#!/bin/bash
cd "$(dirname "$0")/.."
timeout 100 python scripts/kernel_bench.py --only text_qkv --stamps 2>&1 | grep -v globaltimer | tail -4
timeout 100 python scripts/kernel_bench.py --only plain_text_attn_out --stamps 2>&1 | grep -v globaltimer | tail -4
timeout 100 python scripts/kernel_bench.py --only pool_t --stamps 2>&1 | grep -v globaltimer | tail -4
