#!/bin/bash
# Step timeline, second look: how many GEMM launches are active at a time, GPU-wide (needs the -DVB200_STAMPS build).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for fl in 1 2 3 4; do
  timeout 300 python scripts/step_timeline.py --inflight $fl --out $O/s29_timeline_if$fl.json > $O/s29_if$fl.log 2>&1 || tail -5 $O/s29_if$fl.log
done
python - <<PY
import json
for fl in (1, 2, 3, 4):
    j = json.load(open("gpurun_out/r2/s29_timeline_if%d.json" % fl))
    print("inflight", fl, {k: j[k] for k in ("window_ns", "resident_fraction_by_count", "main_loop_fraction_by_count", "gemm_launches_resident_hist", "gemm_launches_in_main_loop_hist", "mean_gemm_ctas_resident")})
PY
