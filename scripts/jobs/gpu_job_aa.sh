#!/bin/bash
mkdir -p gpurun_out
for m in wanda_unstructured dsnot dsnot_elided; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic_$m.csv python bench.py --one-step --method $m > /dev/null 2>&1
  echo "traffic $m: $(wc -l < gpurun_out/traffic_$m.csv) lines"
done
