#!/bin/bash
# 128x192 tiles: op tests, A/B of the step (VB200_BN192=1/0), per-op tables; then the whole -m gpu suite.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "linear_bias or linear_act" > $O/s6_ops.txt 2>&1; echo "exit $?" >> $O/s6_ops.txt
for f in 1 0; do
  VB200_BN192=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --ops-table $O/s6_ops_table_w$f.jsonl > $O/s6_bench_w$f.json 2> $O/s6_bench_w$f.err
  VB200_BN192=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s6_bench_w${f}_if1.json 2>> $O/s6_bench_w$f.err
done
timeout 1500 python -m pytest tests -q -m gpu -x > $O/s6_suite.txt 2>&1; echo "exit $?" >> $O/s6_suite.txt
tail -n 3 $O/s6_ops.txt; tail -n 15 $O/s6_suite.txt
for f in 1 0; do python - <<PY
import json
for suf in ("", "_if1"):
    try:
        j = json.load(open("$O/s6_bench_w$f%s.json" % suf))
        r = j["roofline"]
        print("bn192=$f", suf, round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "gemm TF", round(r["achieved"]), r["families_ms"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
    except Exception as e:
        print("bn192=$f", suf, "ERR", e); print(open("$O/s6_bench_w$f.err").read()[-1500:])
PY
done
python - <<PY
import json
for f in (1, 0):
    print("== bn192", f)
    for l in list(open("$O/s6_ops_table_w%d.jsonl" % f))[:10]: print(l.strip())
PY
