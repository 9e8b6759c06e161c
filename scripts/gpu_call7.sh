#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/select_probe.py > gpurun_out/select_probe.log 2>&1; cat gpurun_out/select_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rowselect_kernel' -s 6 -c 2 -f -o gpurun_out/rowselect_full \
  python scripts/select_probe.py > gpurun_out/ncu_rowselect.log 2>&1
timeout 300 python scripts/chol_probe.py > gpurun_out/chol_probe.log 2>&1; cat gpurun_out/chol_probe.log
