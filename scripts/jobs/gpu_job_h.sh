#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02h_rowselect_probe.log 2>&1; tail -8 gpurun_out/r02h_rowselect_probe.log
timeout 600 python scripts/lora_probe.py > gpurun_out/r02h_lora_probe.log 2>&1; tail -4 gpurun_out/r02h_lora_probe.log
timeout 300 python scripts/dsnot_ncu.py > gpurun_out/r02h_dsnot.log 2>&1; tail -3 gpurun_out/r02h_dsnot.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsnot_walk -c 1 -s 1 -o gpurun_out/r02h_dsnot_walk python scripts/dsnot_ncu.py > gpurun_out/r02h_ncu.log 2>&1; tail -2 gpurun_out/r02h_ncu.log
timeout 900 python -m pytest tests -m gpu -q -k "rowselect or merge or lora or dsnot" 2>&1 | tail -60 > gpurun_out/r02h_pytest_gpu.log; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02h_pytest_gpu.log | tail -20
