#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/obs_sb_probe.py > gpurun_out/r02c_obs_sb_probe.log 2>&1; tail -30 gpurun_out/r02c_obs_sb_probe.log
timeout 900 python -m pytest tests -m gpu -q -k "sparsegpt or obs or calibration_batching" 2>&1 | tail -80 > gpurun_out/r02c_pytest_gpu.log; grep -E "live reference|agreement|passed|failed|Error" gpurun_out/r02c_pytest_gpu.log | tail -20
