#!/bin/bash
# Three CTAs per SM (PCfg MODE 6) for the GEMMs with many tiles: op test (bit-identical), model tests with it on, step A/B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "three_ctas" > $O/s16_ops.txt 2>&1; echo "exit $?" >> $O/s16_ops.txt
tail -n 5 $O/s16_ops.txt
if ! grep -q "exit 0" $O/s16_ops.txt; then exit 1; fi
VB200_TRI=150 timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "golden or shard or task_heads or batch64" > $O/s16_model.txt 2>&1; echo "exit $?" >> $O/s16_model.txt
tail -n 3 $O/s16_model.txt
: > $O/s16_ab.txt
for rep in 1 2; do
for cfg in "0 67" "150 67" "280 67" "150 100" "280 100" "90 67" "150 50"; do
  set -- $cfg
  VB200_TRI=$1 VB200_GRID_PCT=$2 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --ops-table $O/s16_ops_t$1_g$2.jsonl > $O/s16_tmp.json 2> $O/s16_tmp.err
  python - <<PY >> $O/s16_ab.txt
import json
try:
    j = json.load(open("$O/s16_tmp.json")); r = j["roofline"]
    print("rep=$rep tri=$1 grid=$2", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), "full", round(r["achieved_full_grid"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("tri=$1 grid=$2 ERR", e, open("$O/s16_tmp.err").read()[-600:])
PY
done
done
cat $O/s16_ab.txt
head -6 $O/s16_ops_t150_g67.jsonl; head -6 $O/s16_ops_t0_g67.jsonl
