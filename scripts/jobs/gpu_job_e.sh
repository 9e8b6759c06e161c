#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02e_rowselect_probe.log 2>&1; tail -12 gpurun_out/r02e_rowselect_probe.log
timeout 900 python -m pytest tests -m gpu -q -k "rowselect or dsnot or composite or wanda or batching or shared" 2>&1 | tail -60 > gpurun_out/r02e_pytest_gpu.log; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02e_pytest_gpu.log | tail -30
