#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02d_rowselect_probe.log 2>&1; tail -12 gpurun_out/r02d_rowselect_probe.log
timeout 600 python scripts/obs_sb_probe.py > gpurun_out/r02d_obs_sb_probe.log 2>&1; grep "C=11008\|R=11008" gpurun_out/r02d_obs_sb_probe.log | tail -24
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02d_pytest_gpu.log; grep -E "live reference|passed|failed|^FAILED|^E  " gpurun_out/r02d_pytest_gpu.log | tail -30
