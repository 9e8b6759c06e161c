#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/lora_probe.py > gpurun_out/r02i_lora_probe.log 2>&1; tail -4 gpurun_out/r02i_lora_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowselect_cta -c 1 -s 1 -o gpurun_out/r02i_rowselect_cta python scripts/rowselect_ncu.py > gpurun_out/r02i_ncu.log 2>&1; tail -2 gpurun_out/r02i_ncu.log
timeout 600 python -m pytest tests -m gpu -q -k "lora or merge" 2>&1 | tail -5
