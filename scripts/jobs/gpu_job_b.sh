#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02b_pytest_gpu.log; tail -25 gpurun_out/r02b_pytest_gpu.log
timeout 600 python scripts/obs_sb_probe.py > gpurun_out/r02b_obs_sb_probe.log 2>&1; tail -30 gpurun_out/r02b_obs_sb_probe.log
