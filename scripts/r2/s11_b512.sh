#!/bin/bash
# Batch 512: persistent-grid size x LayerNorm fold x CTA-pair selection; batch 64: balanced grid A/B; chained-FFN model test.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "chained_ffn" > $O/s11_chain_test.txt 2>&1; echo "exit $?" >> $O/s11_chain_test.txt
: > $O/s11_ab.txt
line() {
python - "$@" <<'PY' >> gpurun_out/r2/s11_ab.txt
import json, sys
tag = sys.argv[1]
try:
    j = json.load(open("gpurun_out/r2/s11_tmp.json")); r = j["roofline"]
    print(tag, round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), "full", round(r["achieved_full_grid"]), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print(tag, "ERR", e, open("gpurun_out/r2/s11_tmp.err").read()[-600:])
PY
}
for rep in 1 2; do
  for bal in 1 0; do
    VB200_GRID_BALANCE=$bal timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dtype fp16 > $O/s11_tmp.json 2> $O/s11_tmp.err
    line "b64 rep=$rep balance=$bal"
  done
done
for cfg in "67 0 auto" "100 0 auto" "100 1 auto" "67 1 auto" "100 0 0" "100 1 0" "85 0 auto"; do
  set -- $cfg
  VB200_GRID_PCT=$1 VB200_LNFOLD=$2 VB200_PAIR=$3 timeout 300 python bench.py --batch 512 --steps 30 --warmup 4 --no-cpu-baseline --dtype fp16 --ops-table $O/s11_ops_b512_g$1_f$2_p$3.jsonl > $O/s11_tmp.json 2> $O/s11_tmp.err
  line "b512 grid=$1 fold=$2 pair=$3"
done
cat $O/s11_chain_test.txt | tail -4; cat $O/s11_ab.txt
