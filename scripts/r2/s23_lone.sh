#!/bin/bash
# Small-batch latency: one CTA per SM with a 6-stage ring (PCfg MODE 7, default for forwards of <= 320 rows) vs the default kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "three_ctas" > $O/s23_ops.txt 2>&1; echo "exit $?" >> $O/s23_ops.txt
tail -n 3 $O/s23_ops.txt
if ! grep -q "exit 0" $O/s23_ops.txt; then exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_round2.py -x -q -m gpu > $O/s23_model.txt 2>&1; echo "exit $?" >> $O/s23_model.txt
tail -n 3 $O/s23_model.txt
: > $O/s23_ab.txt
for rep in 1 2; do
for b in 1 4 8; do
for lone in 320 0; do
  VB200_LONE_ROWS=$lone timeout 300 python bench.py --batch $b --inflight 1 --steps 300 --warmup 10 --no-cpu-baseline --dtype fp16 > $O/s23_tmp.json 2> $O/s23_tmp.err
  python - <<PY >> $O/s23_ab.txt
import json
try:
    j = json.load(open("$O/s23_tmp.json")); r = j["roofline"]
    print("rep=$rep batch=$b lone_rows=$lone", "ms/forward", round(j["ms_per_step"], 4), "pairs/s", round(j["value"]), "e2e ms", round(j["e2e"]["ms_per_step"], 4), r["families_ms"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("b=$b lone=$lone ERR", e, open("$O/s23_tmp.err").read()[-600:])
PY
done
done
done
cat $O/s23_ab.txt
