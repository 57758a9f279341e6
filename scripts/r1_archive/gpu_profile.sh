#!/bin/bash
# Launch list (per-kernel device time) of the bench step + CPU thread-count probe.  Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?" > gpurun_out/status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --pdl on > gpurun_out/bench_pdl.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 512 > gpurun_out/bench_b512.log 2>&1
timeout 600 python scripts/cpu_threads_probe.py > gpurun_out/cpu_threads.log 2>&1
tail -1 gpurun_out/bench_plain.log; tail -1 gpurun_out/bench_pdl.log; tail -1 gpurun_out/bench_nograph.log; tail -1 gpurun_out/bench_b512.log; cat gpurun_out/cpu_threads.log
