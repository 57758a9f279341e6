#!/bin/bash
# Round 2, GPU session 3: LayerNorm fold -- op chain test, model tests, A/B bench (fold on / off), per-op table.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "fold_chain" > $O/s3_ops.txt 2>&1; echo "exit $?" >> $O/s3_ops.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "fold or cached or fp32x_tiny" > $O/s3_round2.txt 2>&1; echo "exit $?" >> $O/s3_round2.txt
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/s3_model.txt 2>&1; echo "exit $?" >> $O/s3_model.txt
for f in 1 0; do
  VB200_LNFOLD=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --ops-table $O/s3_ops_table_fold$f.jsonl > $O/s3_bench_fold$f.json 2> $O/s3_bench_fold$f.err
  VB200_LNFOLD=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s3_bench_fold${f}_if1.json 2>> $O/s3_bench_fold$f.err
done
tail -n 5 $O/s3_ops.txt; tail -n 25 $O/s3_round2.txt; tail -n 25 $O/s3_model.txt
for f in 1 0; do python - <<PY
import json
for suf in ("", "_if1"):
    try:
        j = json.load(open("$O/s3_bench_fold$f%s.json" % suf))
        r = j["roofline"]
        print("fold=$f", suf, round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), "launches", j["launches_per_step"], "gemm TF", round(r["achieved"]), r["families_ms"], "parity", j["parity"]["max_abs_err_vs_fp32_oracle"])
    except Exception as e:
        print("fold=$f", suf, "ERR", e); print(open("$O/s3_bench_fold$f.err").read()[-1500:])
PY
done
