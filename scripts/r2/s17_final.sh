#!/bin/bash
# Round-2 final evidence on one B200: the whole -m gpu suite, smoke(), the bench lines, the ncu launch list and ncu --set full
# captures of the dominant GEMMs (raw page exported as CSV; the .ncu-rep files stay in gpurun_out/).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/s17_env.txt
timeout 1800 python -m pytest tests -q -m gpu > $O/s17_suite.txt 2>&1; echo "exit $?" >> $O/s17_suite.txt
timeout 300 python __graft_entry__.py smoke > $O/s17_smoke.txt 2>&1; echo "exit $?" >> $O/s17_smoke.txt
timeout 400 python bench.py --steps 40 --warmup 5 --ops-table $O/s17_ops_table.jsonl > $O/s17_bench_1gpu.json 2> $O/s17_bench_1gpu.err
timeout 400 python bench.py > $O/s17_bench_default_flags.json 2> $O/s17_bench_default_flags.err
timeout 400 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/s17_bench_200steps.json 2> $O/s17_bench_200.err
timeout 300 python bench.py --steps 200 --warmup 5 --inflight 1 --no-cpu-baseline > $O/s17_bench_1gpu_inflight1.json 2> $O/s17_bench_if1.err
timeout 300 python bench.py --steps 40 --warmup 5 --batch 512 --no-cpu-baseline --ops-table $O/s17_ops_table_b512.jsonl > $O/s17_bench_b512.json 2> $O/s17_bench_b512.err
timeout 300 python bench.py --steps 50 --warmup 5 --dtype fp32x --no-cpu-baseline > $O/s17_bench_fp32x.json 2> $O/s17_bench_fp32x.err
timeout 300 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline --dtype fp16 > $O/s17_bench_sustained.json 2> $O/s17_bench_sus.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 2 > $O/s17_bench_reference.json 2> $O/s17_bench_ref.err
# ncu: launch list of one short bench run, then full captures: the GELU FFN-in GEMM and three consecutive GEMMs of a forward
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s17_launches_all.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s17_bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/s17_launches_all.csv $O/s17_launches > $O/s17_launch_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_persistent_kernel<128, 0, 1, 1, 0>" -s 40 -c 2 -f -o $O/s17_gemm_gelu \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s17_ncu_full_gelu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_persistent_kernel -s 400 -c 3 -f -o $O/s17_gemm \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 --inflight 1 > $O/s17_ncu_full.log 2>&1
ncu -i $O/s17_gemm.ncu-rep --page raw --csv > $O/s17_gemm.raw.csv 2>/dev/null
ncu -i $O/s17_gemm_gelu.ncu-rep --page raw --csv > $O/s17_gemm_gelu.raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/s17_gemm_gelu.raw.csv $O/s17_traffic.json > $O/s17_traffic.txt 2>&1 || python scripts/ncu_traffic.py $O/s17_gemm.raw.csv $O/s17_traffic.json > $O/s17_traffic.txt 2>&1
rm -f $O/s17_launches_all.csv.tmp
tail -n 4 $O/s17_suite.txt; tail -n 2 $O/s17_smoke.txt
for f in 1gpu default_flags 200steps 1gpu_inflight1 b512 fp32x sustained reference; do echo "== $f"; cut -c1-600 $O/s17_bench_$f.json; done
cut -c1-600 $O/s17_launch_summary.txt; cut -c1-400 $O/s17_traffic.txt; ls -la $O/s17_gemm*
