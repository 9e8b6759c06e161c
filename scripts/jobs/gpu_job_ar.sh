#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "lora" 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lora_linear -s 2 -c 1 -o gpurun_out/lora_linear_full -f python scripts/lora_linear_once.py > gpurun_out/ncu_ll.log 2>&1
tail -3 gpurun_out/ncu_ll.log
