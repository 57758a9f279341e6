#!/bin/bash
# Large batch: cluster-LayerNorm GEMM epilogue (fused_layernorm) vs GEMM + HBM-bound row kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s22_ab.txt
for rep in 1 2; do
for cfg in "512 0" "512 1" "256 0" "256 1"; do
  set -- $cfg
  FL=""; if [ "$2" = "1" ]; then FL="--fused-ln"; fi
  timeout 300 python bench.py --batch $1 --steps 30 --warmup 4 --no-cpu-baseline --dtype fp16 $FL --ops-table $O/s22_ops_b$1_f$2.jsonl > $O/s22_tmp.json 2> $O/s22_tmp.err
  python - <<PY >> $O/s22_ab.txt
import json
try:
    j = json.load(open("$O/s22_tmp.json")); r = j["roofline"]
    print("rep=$rep batch=$1 fused_ln=$2", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), r["families_ms"], j["clocks"]["sm_mhz"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("b=$1 f=$2 ERR", e, open("$O/s22_tmp.err").read()[-600:])
PY
done
done
cat $O/s22_ab.txt; head -8 $O/s22_ops_b512_f1.jsonl
