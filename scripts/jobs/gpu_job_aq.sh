#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "lora" 2>&1 | tail -15
timeout 300 python scripts/lora_linear_probe.py 2>&1 | tee gpurun_out/r02ar_lora_linear_probe.log | tail -40
