#!/bin/bash
# LayerNorm fold A/B (fold on / off), one and two batches in flight, with the per-op table; plus the fold op test.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
T=${TAG:-s4}
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "fold_chain or linear_bias" > $O/${T}_ops.txt 2>&1; echo "exit $?" >> $O/${T}_ops.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "fold or cached" > $O/${T}_round2.txt 2>&1; echo "exit $?" >> $O/${T}_round2.txt
for f in 1 0; do
  VB200_LNFOLD=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --ops-table $O/${T}_ops_table_fold$f.jsonl > $O/${T}_bench_fold$f.json 2> $O/${T}_bench_fold$f.err
  VB200_LNFOLD=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/${T}_bench_fold${f}_if1.json 2>> $O/${T}_bench_fold$f.err
done
tail -n 3 $O/${T}_ops.txt; tail -n 6 $O/${T}_round2.txt
for f in 1 0; do python - <<PY
import json
for suf in ("", "_if1"):
    try:
        j = json.load(open("$O/${T}_bench_fold$f%s.json" % suf))
        r = j["roofline"]
        print("fold=$f", suf, round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), r["families_ms"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
    except Exception as e:
        print("fold=$f", suf, "ERR", e); print(open("$O/${T}_bench_fold$f.err").read()[-1500:])
PY
done
python - <<PY
import json
for f in (1, 0):
    print("== fold", f)
    for l in list(open("$O/${T}_ops_table_fold%d.jsonl" % f))[:12]: print(l.strip())
PY
