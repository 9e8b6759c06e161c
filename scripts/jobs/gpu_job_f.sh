#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/rowselect_probe.py > gpurun_out/r02f_rowselect_probe.log 2>&1; tail -8 gpurun_out/r02f_rowselect_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowselect_cta -c 1 -s 1 -o gpurun_out/r02f_rowselect_cta python scripts/rowselect_ncu.py > gpurun_out/r02f_ncu.log 2>&1; tail -3 gpurun_out/r02f_ncu.log
