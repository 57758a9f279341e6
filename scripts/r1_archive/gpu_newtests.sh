#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_tasks.py tests/test_gpu_library_bar.py -x -q -m gpu --durations=5 2>&1 | tail -14
grep "library_bar\|B512\|1000x1000" gpurun_out/parity.jsonl
