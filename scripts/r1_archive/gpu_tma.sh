#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -6
timeout 100 python scripts/kernel_bench.py --only text_qkv --stamps 2>&1 | grep -v globaltimer | tail -4
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("$name ERR", e)
PY
}
run tma1 A=1
run tma0 VB200_TMASTORE=0
run tma1b A=1
run tma0b VB200_TMASTORE=0
