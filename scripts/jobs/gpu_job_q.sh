#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sgpt_step_probe.py > gpurun_out/r02q_sgpt_step_probe.log 2>&1; cat gpurun_out/r02q_sgpt_step_probe.log | tail -12
echo "--- CUDA_DEVICE_MAX_CONNECTIONS=32"
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python scripts/sgpt_step_probe.py > gpurun_out/r02q_sgpt_step_probe_mc32.log 2>&1; cat gpurun_out/r02q_sgpt_step_probe_mc32.log | tail -12
timeout 600 python -m pytest tests -m gpu -q -x -k "rowselect" 2>&1 | tail -3
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02q_rowselect_probe.log 2>&1; tail -8 gpurun_out/r02q_rowselect_probe.log
