#!/bin/bash
mkdir -p gpurun_out
for m in wanda_nm; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic_$m.csv python bench.py --one-step --method $m > /dev/null 2>&1
  echo "traffic $m: $(wc -l < gpurun_out/traffic_$m.csv) lines"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:colstats_batch -c 1 -o gpurun_out/r02ak_colstats_batch python bench.py --one-step --method wanda_nm > gpurun_out/r02ak_ncu.log 2>&1; tail -1 gpurun_out/r02ak_ncu.log
