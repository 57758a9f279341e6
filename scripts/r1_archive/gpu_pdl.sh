#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out

timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "pdl or graph" 2>&1 | tail -2
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/nopdl.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --pdl on > gpurun_out/pdl.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --pdl on --inflight 1 > gpurun_out/pdl_if1.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --pdl on --inflight 3 > gpurun_out/pdl_if3.log 2>&1
python - <<'PY'
import json
for n in ("nopdl", "pdl", "pdl_if1", "pdl_if3"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
