#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowselect_cta -c 1 -s 1 -o gpurun_out/r02ae_rowselect python scripts/rowselect_ncu.py > gpurun_out/r02ae_ncu.log 2>&1; tail -2 gpurun_out/r02ae_ncu.log
