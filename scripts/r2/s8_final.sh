#!/bin/bash
# Round-2 evidence refresh on one B200: the whole -m gpu suite, smoke(), the bench lines, the ncu launch list and one
# ncu --set full capture of the dominant GEMM (raw page exported as CSV; the .ncu-rep stays in gpurun_out/).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/s8_env.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/s8_suite.txt 2>&1; echo "exit $?" >> $O/s8_suite.txt
timeout 300 python __graft_entry__.py smoke > $O/s8_smoke.txt 2>&1; echo "exit $?" >> $O/s8_smoke.txt
timeout 400 python bench.py --steps 200 --warmup 5 --ops-table $O/s8_ops_table.jsonl > $O/s8_bench_1gpu.json 2> $O/s8_bench_1gpu.err
timeout 300 python bench.py --steps 200 --warmup 5 --inflight 1 --no-cpu-baseline > $O/s8_bench_1gpu_inflight1.json 2> $O/s8_bench_if1.err
timeout 300 python bench.py --steps 40 --warmup 5 --batch 512 --no-cpu-baseline > $O/s8_bench_b512.json 2> $O/s8_bench_b512.err
timeout 300 python bench.py --steps 50 --warmup 5 --dtype fp32x --no-cpu-baseline > $O/s8_bench_fp32x.json 2> $O/s8_bench_fp32x.err
timeout 300 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline --dtype fp16 > $O/s8_bench_sustained.json 2> $O/s8_bench_sus.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 2 > $O/s8_bench_reference.json 2> $O/s8_bench_ref.err
# ncu: launch list of one short bench run (after warm-up), then one full capture of the largest GEMM of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s8_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s8_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_persistent_kernel -s 400 -c 3 -f -o $O/s8_gemm \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s8_ncu_full.log 2>&1
ncu -i $O/s8_gemm.ncu-rep --page raw --csv > $O/s8_gemm.raw.csv 2>/dev/null
tail -n 4 $O/s8_suite.txt; tail -n 2 $O/s8_smoke.txt
for f in 1gpu 1gpu_inflight1 b512 fp32x sustained reference; do echo "== $f"; cut -c1-700 $O/s8_bench_$f.json; done
wc -l $O/s8_launches.csv; ls -la $O/s8_gemm*
