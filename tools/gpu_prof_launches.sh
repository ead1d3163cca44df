#!/bin/bash
# launch list of the bench command (ncu --metrics gpu__time_duration.sum), after the same command exited 0 without ncu
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_short.json 2>/dev/null || exit 1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu3.log 2>&1
wc -l gpurun_out/launches_bench.csv; du -sh gpurun_out
