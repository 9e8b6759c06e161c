#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "dsnot" 2>&1 | grep -E "walk2|passed|failed|Error|error|assert" | head -40 > gpurun_out/r02r_pytest_gpu.log; cat gpurun_out/r02r_pytest_gpu.log
for C in 4096 11008; do timeout 300 python scripts/dsnot_ncu.py $C >> gpurun_out/r02r_dsnot.log 2>&1; done; cat gpurun_out/r02r_dsnot.log
