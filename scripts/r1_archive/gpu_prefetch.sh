#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_tasks.py tests/test_gpu_library_bar.py -x -q -m gpu --durations=4 2>&1 | tail -10
grep "library_bar" gpurun_out/parity.jsonl | tail -1
for pf in 1 0; do
for nf in 2 1; do
VB200_PREFETCH=$pf timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --inflight $nf > gpurun_out/pf${pf}_if$nf.log 2>&1
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/pf${pf}_if$nf.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("prefetch $pf inflight $nf", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("prefetch $pf ERR", e)
PY
done
done
