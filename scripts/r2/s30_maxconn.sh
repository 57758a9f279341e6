#!/bin/bash
# Do more hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) let more of the in-flight graphs' branches run concurrently?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
: > $O/s30_ab.txt
for rep in 1 2; do
for mc in 8 32; do
for fl in 2 3 4; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight $fl > $O/s30_tmp.json 2> $O/s30_tmp.err
  python - <<PY >> $O/s30_ab.txt
import json
try:
    j = json.load(open("$O/s30_tmp.json"))
    print("rep=$rep max_connections=$mc inflight=$fl", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("mc=$mc fl=$fl ERR", e, open("$O/s30_tmp.err").read()[-400:])
PY
done
done
done
cat $O/s30_ab.txt
