#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sgpt_step_probe.py > gpurun_out/r02p_sgpt_step_probe.log 2>&1; cat gpurun_out/r02p_sgpt_step_probe.log | tail -12
timeout 600 python -m pytest tests -m gpu -q -x -k "rowselect or wanda or dsnot" 2>&1 | tail -15 > gpurun_out/r02p_pytest_gpu.log; tail -5 gpurun_out/r02p_pytest_gpu.log
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02p_rowselect_probe.log 2>&1; tail -8 gpurun_out/r02p_rowselect_probe.log
