#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; "$@" > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"])
except Exception as e:
    print("$name ERR", e)
PY
}
run base timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline
run pdl timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --pdl on
run b512auto timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 512
run b256auto timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 256
