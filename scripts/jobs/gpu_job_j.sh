#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02j_rowselect_probe.log 2>&1; tail -8 gpurun_out/r02j_rowselect_probe.log
timeout 600 python scripts/lora_probe.py > gpurun_out/r02j_lora_probe.log 2>&1; tail -4 gpurun_out/r02j_lora_probe.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -100 > gpurun_out/r02j_pytest_gpu.log; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02j_pytest_gpu.log | tail -20
