#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rowselect or wanda" 2>&1 | tail -3
timeout 300 python scripts/rowselect_probe.py > gpurun_out/r02af_rowselect_probe.log 2>&1; tail -7 gpurun_out/r02af_rowselect_probe.log
