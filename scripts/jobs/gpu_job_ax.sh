#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/chol_sp_probe.py 4096 11008 2>&1 | tee gpurun_out/r02ax_chol_sp_probe.log | tail -14
timeout 600 python -m pytest tests -m gpu -q -x -k "chol or sparsegpt" 2>&1 | tail -4
