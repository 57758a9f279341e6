#!/bin/bash
# Chained FFN launch (gemm_chain.cu): op test (bit-identical to two launches), model tests, step A/B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm_chain" > $O/s10_ops.txt 2>&1; echo "exit $?" >> $O/s10_ops.txt
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny_model_all_outputs or golden or shard or task_heads" > $O/s10_model.txt 2>&1; echo "exit $?" >> $O/s10_model.txt
: > $O/s10_ab.txt
for rep in 1 2; do
for cfg in "1 67" "0 67" "1 100" "1 50"; do
  set -- $cfg
  for fl in 2 1; do
    VB200_CHAIN=$1 VB200_CHAIN_GRID_PCT=$2 timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $fl --ops-table $O/s10_ops_table_c$1_g$2.jsonl > $O/s10_tmp.json 2> $O/s10_tmp.err
    python - <<PY >> $O/s10_ab.txt
import json
try:
    j = json.load(open("$O/s10_tmp.json")); r = j["roofline"]
    print("rep=$rep chain=$1 chain_grid=$2 inflight=$fl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), "full", round(r["achieved_full_grid"]), r["families_ms"], j["clocks"]["sm_mhz"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("chain=$1 grid=$2 inflight=$fl ERR", e, open("$O/s10_tmp.err").read()[-600:])
PY
  done
done
done
tail -n 4 $O/s10_ops.txt; tail -n 6 $O/s10_model.txt; cat $O/s10_ab.txt
head -8 $O/s10_ops_table_c1_g67.jsonl
