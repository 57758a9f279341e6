#!/bin/bash
# Batch 512: how many waves of 256x256 pair tiles a GEMM needs before the CTA-pair kernel takes it; + the GELU GEMM under ncu --set full;
# + the timeline hook test.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "timeline or chained" > $O/s19_tests.txt 2>&1; echo "exit $?" >> $O/s19_tests.txt
: > $O/s19_ab.txt
for rep in 1 2; do
for w in 4 3 2 1; do
  VB200_PAIR_MIN_WAVES=$w timeout 300 python bench.py --batch 512 --steps 30 --warmup 4 --no-cpu-baseline --dtype fp16 --ops-table $O/s19_ops_w$w.jsonl > $O/s19_tmp.json 2> $O/s19_tmp.err
  python - <<PY >> $O/s19_ab.txt
import json
try:
    j = json.load(open("$O/s19_tmp.json")); r = j["roofline"]
    print("rep=$rep pair_min_waves=$w", round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), "frac", round(r["frac"], 3), r["families_ms"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
except Exception as e:
    print("w=$w ERR", e, open("$O/s19_tmp.err").read()[-600:])
PY
done
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_persistent_kernel<\(int\)128, \(bool\)0, \(int\)1" -s 40 -c 2 -f -o $O/s19_gemm_gelu \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s19_ncu_full_gelu.log 2>&1
ncu -i $O/s19_gemm_gelu.ncu-rep --page raw --csv > $O/s19_gemm_gelu.raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/s19_gemm_gelu.raw.csv $O/s19_traffic_gelu.json > $O/s19_traffic.txt 2>&1
tail -n 3 $O/s19_tests.txt; cat $O/s19_ab.txt; cut -c1-900 $O/s19_traffic.txt; for w in 4 3 2 1; do echo "w=$w"; head -9 $O/s19_ops_w$w.jsonl; done
