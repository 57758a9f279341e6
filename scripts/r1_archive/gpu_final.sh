#!/bin/bash
# evidence refresh: full gpu suite, bench lines, reference arm, smoke, launch list, one ncu --set full of the GELU GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/final_1gpu.log 2>gpurun_out/final_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 40 --warmup 5 --inflight 1 --no-cpu-baseline > gpurun_out/final_if1.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --batch 512 --no-cpu-baseline > gpurun_out/final_b512.log 2>&1
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/final_ref.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:'gemm_persistent' -s 30 -c 1 \
     -o gpurun_out/prof_gemm_gelu -f python scripts/kernel_bench.py --only text_ffn_in_gelu --reps 20 > gpurun_out/ncu_gelu.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for n in ("final_1gpu", "final_if1", "final_b512", "final_ref"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
        r = j.get("roofline", {})
        print(n, round(j["value"], 1), round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 1), "gemm TF", r.get("achieved"), r.get("frac"), r.get("whole_step_frac"), j.get("clocks"), j.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
