#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VB200_WIDE2=1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "persistent and (linear_bias or linear_act)" 2>&1 | tail -2
VB200_BN=256 timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "full_model_task_heads or batch64" 2>&1 | tail -2
for b in 64 512; do
  echo "== batch $b"
  VB200_WIDE2=0 timeout 100 python scripts/kernel_bench.py --batch $b --only img_qkv --stamps 2>&1 | grep -v "globaltimer\|epilogue of" | tail -3
  timeout 100 python scripts/kernel_bench.py --batch $b --bn 256 --only img_qkv --stamps 2>&1 | grep -v "globaltimer\|epilogue of" | tail -3
done
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/$name.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), r["families_ms"], round(r["achieved"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("$name ERR", e)
PY
}
run base VB200_WIDE2=0
run wide2 VB200_WIDE2=1 VB200_BN=256
run base_b VB200_WIDE2=0
run wide2_b VB200_WIDE2=1 VB200_BN=256
