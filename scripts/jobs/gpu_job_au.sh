#!/bin/bash
bash scripts/gpu_round_check.sh r02final3
for M in wanda_nm wanda_unstructured sparsegpt dsnot dsnot_elided; do
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic_$M.csv python bench.py --one-step --method $M > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:colstats_batch -c 1 -o gpurun_out/colstats_batch_waves_full -f python bench.py --one-step --method wanda_nm > /dev/null 2>&1
ls -la gpurun_out/traffic_*.csv gpurun_out/colstats_batch_waves_full.ncu-rep
