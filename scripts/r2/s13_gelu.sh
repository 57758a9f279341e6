#!/bin/bash
# 11-slot GELU (x * sat(0.5 + x Q(x^2))) and the full grid from two waves up: op tests, chained-FFN model test, step A/B at batch 64 / 512.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/s13_ops.txt 2>&1; echo "exit $?" >> $O/s13_ops.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "chained_ffn or fp32x" > $O/s13_chain_test.txt 2>&1; echo "exit $?" >> $O/s13_chain_test.txt
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/s13_model.txt 2>&1; echo "exit $?" >> $O/s13_model.txt
: > $O/s13_ab.txt
line() {
python - "$@" <<'PY' >> gpurun_out/r2/s13_ab.txt
import json, sys
tag = sys.argv[1]
try:
    j = json.load(open("gpurun_out/r2/s13_tmp.json")); r = j["roofline"]
    print(tag, round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), "full", round(r["achieved_full_grid"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print(tag, "ERR", e, open("gpurun_out/r2/s13_tmp.err").read()[-600:])
PY
}
for rep in 1 2 3; do
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dtype fp16 --ops-table $O/s13_ops_b64.jsonl > $O/s13_tmp.json 2> $O/s13_tmp.err
  line "b64 fp16 rep=$rep"
done
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $O/s13_tmp.json 2> $O/s13_tmp.err; line "b64 default(bf16) 40 steps"; cp $O/s13_tmp.json $O/s13_bench_default.json
timeout 300 python bench.py --batch 512 --steps 30 --warmup 4 --no-cpu-baseline --dtype fp16 --ops-table $O/s13_ops_b512.jsonl > $O/s13_tmp.json 2> $O/s13_tmp.err; line "b512 fp16"
timeout 300 python bench.py --batch 512 --steps 30 --warmup 4 --no-cpu-baseline > $O/s13_tmp.json 2> $O/s13_tmp.err; line "b512 default(bf16)"; cp $O/s13_tmp.json $O/s13_bench_b512.json
tail -n 3 $O/s13_ops.txt; tail -n 3 $O/s13_chain_test.txt; tail -n 3 $O/s13_model.txt; cat $O/s13_ab.txt; head -3 $O/s13_ops_b64.jsonl; head -3 $O/s13_ops_b512.jsonl
