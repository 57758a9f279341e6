#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_tasks.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/e$rep.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/e$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("run $rep", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("run $rep ERR", e)
PY
done
VB200_PDL=full timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
VB200_PDL=full timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inflight 1 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inflight 1 2>&1 | tail -1 | cut -c1-200
