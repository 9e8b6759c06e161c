#!/bin/bash
# One gpurun call: GPU parity suite, the default bench line, Cholesky / OBS probe, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.err
timeout 300 python scripts/chol_probe.py > gpurun_out/chol_probe.log 2>&1
cat gpurun_out/chol_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_wanda_nm.csv \
  python bench.py --steps 2 --warmup 3 --no-other-methods --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'colstats|nm_kernel' -c 14 -f -o gpurun_out/wanda_full \
  python bench.py --steps 2 --warmup 3 --no-other-methods --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
