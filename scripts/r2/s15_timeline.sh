#!/bin/bash
# Per-SM GEMM-CTA occupancy of the step (scripts/step_timeline.py; needs the -DVB200_STAMPS build).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for fl in 2 1 3; do
  timeout 300 python scripts/step_timeline.py --inflight $fl --out $O/s15_timeline_if$fl.json > $O/s15_if$fl.log 2>&1 || tail -5 $O/s15_if$fl.log
done
VB200_GRID_PCT=100 timeout 300 python scripts/step_timeline.py --inflight 2 --out $O/s15_timeline_if2_g100.json > $O/s15_if2_g100.log 2>&1
VB200_GRID_PCT=33 timeout 300 python scripts/step_timeline.py --inflight 3 --out $O/s15_timeline_if3_g33.json > $O/s15_if3_g33.log 2>&1
for f in $O/s15_timeline_*.json; do echo $f; python - <<PY
import json
j = json.load(open("$f"))
print({k: j[k] for k in j if k not in ("families", "config")})
for r in j.get("families", [])[:12]: print(r)
PY
done
