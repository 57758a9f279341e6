#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VB200_DEEP_MAXM=128 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/deep_m128.log 2>&1
VB200_DEEP_MINKB=32 VB200_DEEP_MAXM=128 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/deep_kb32.log 2>&1
VB200_DEEP=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --inflight 3 > gpurun_out/nodeep_if3.log 2>&1
python - <<'PY'
import json
for n in ("deep_m128", "deep_kb32", "nodeep_if3"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
