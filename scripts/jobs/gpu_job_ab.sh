#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stats or rowselect or wanda" 2>&1 | tail -3
timeout 300 python scripts/rowselect_probe.py > gpurun_out/r02ab_rowselect_probe.log 2>&1; tail -7 gpurun_out/r02ab_rowselect_probe.log
for C in 4096 11008; do timeout 300 python scripts/dsnot_stats_ncu.py $C 2>&1 | tail -1; done | tee gpurun_out/r02ab_dsnot_stats.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:colstats -c 1 -s 2 -o gpurun_out/r02ab_dsnot_stats python scripts/dsnot_stats_ncu.py 4096 > gpurun_out/r02ab_ncu1.log 2>&1; tail -1 gpurun_out/r02ab_ncu1.log
for C in 4096 11008; do
  timeout 900 ncu --set full --clock-control none -k regex:hessian_syrk2 -c 1 -s 1 -o gpurun_out/r02ab_hessian_C$C python scripts/hessian_ncu.py $C > gpurun_out/r02ab_ncu_h$C.log 2>&1; tail -1 gpurun_out/r02ab_ncu_h$C.log
done
