#!/bin/bash
# Small-batch latency: cluster-LayerNorm epilogue (151 launches) vs split (213); lone-CTA threshold at batch 16 / 32.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s24_ab.txt
run() {
  tag="$1"; shift
  timeout 300 python bench.py --inflight 1 --steps 300 --warmup 10 --no-cpu-baseline --dtype fp16 "$@" > $O/s24_tmp.json 2> $O/s24_tmp.err
  python - "$tag" <<'PY' >> gpurun_out/r2/s24_ab.txt
import json, sys
try:
    j = json.load(open("gpurun_out/r2/s24_tmp.json")); r = j["roofline"]
    print(sys.argv[1], "ms/forward", round(j["ms_per_step"], 4), "pairs/s", round(j["value"]), "e2e ms", round(j["e2e"]["ms_per_step"], 4), "launches", j["launches_per_step"], r["families_ms"])
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/r2/s24_tmp.err").read()[-600:])
PY
}
for rep in 1 2; do
  run "rep=$rep b=1 split" --batch 1
  run "rep=$rep b=1 fused_ln" --batch 1 --fused-ln
  run "rep=$rep b=8 split" --batch 8
  run "rep=$rep b=8 fused_ln" --batch 8 --fused-ln
  VB200_LONE_ROWS=320 run "rep=$rep b=16 lone_rows=320(off here)" --batch 16
  VB200_LONE_ROWS=640 run "rep=$rep b=16 lone_rows=640" --batch 16
  VB200_LONE_ROWS=320 run "rep=$rep b=32 lone off" --batch 32
  VB200_LONE_ROWS=1280 run "rep=$rep b=32 lone_rows=1280" --batch 32
done
cat $O/s24_ab.txt
