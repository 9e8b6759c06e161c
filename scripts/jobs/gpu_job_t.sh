#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "dsnot" 2>&1 | grep -E "walk2|passed|failed|Error|error|assert" | head -40 > gpurun_out/r02t_pytest_gpu.log; cat gpurun_out/r02t_pytest_gpu.log
rm -f gpurun_out/r02t_dsnot.log
for C in 4096 11008; do timeout 300 python scripts/dsnot_ncu.py $C >> gpurun_out/r02t_dsnot.log 2>&1; done; cat gpurun_out/r02t_dsnot.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsnot_walk2 -c 1 -s 1 -o gpurun_out/r02t_dsnot_walk2 python scripts/dsnot_ncu.py > gpurun_out/r02t_ncu.log 2>&1; tail -1 gpurun_out/r02t_ncu.log
