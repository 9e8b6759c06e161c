#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vlmc -c 4000 --csv \
  --log-file gpurun_out/launches_default_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
wc -l gpurun_out/launches_default_bench.csv
