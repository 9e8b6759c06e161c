#!/bin/bash
# ncu --set full captures of the dominant kernels (one GPU): raw pages are exported to profiles/ from the .ncu-rep files
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'colstats_kernel|nm_batch_kernel' -c 8 -f \
  -o gpurun_out/wanda_step_full python bench.py --steps 2 --warmup 3 --no-other-methods --no-cpu-baseline --no-graph > gpurun_out/ncu_wanda.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'hessian_syrk2_kernel' -s 6 -c 2 -f \
  -o gpurun_out/hessian_pair_full python scripts/hessian_pair_probe.py > gpurun_out/ncu_hessian.log 2>&1
ls -la gpurun_out/*.ncu-rep
