#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stats or dsnot" 2>&1 | tail -3
for C in 4096 11008; do timeout 300 python scripts/dsnot_stats_ncu.py $C 2>&1 | tail -1; done | tee gpurun_out/r02ac_dsnot_stats.log
