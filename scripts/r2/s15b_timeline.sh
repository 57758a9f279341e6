#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python scripts/step_timeline.py --inflight 2 --out $O/s15b_timeline_if2.json > $O/s15b_if2.log 2>&1 || tail -5 $O/s15b_if2.log
timeout 300 python scripts/step_timeline.py --inflight 1 --out $O/s15b_timeline_if1.json > $O/s15b_if1.log 2>&1 || tail -5 $O/s15b_if1.log
python - <<PY
import json
for f in ("s15b_timeline_if2", "s15b_timeline_if1"):
    j = json.load(open("gpurun_out/r2/%s.json" % f))
    print(f, {k: j[k] for k in j if k not in ("families", "config", "one_tile_cta_cycles")})
    for r in j["one_tile_cta_cycles"][:8]: print(r)
PY
