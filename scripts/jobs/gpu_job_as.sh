#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/chol_probe.py 4096 11008 2>&1 | tee gpurun_out/r02as_chol_probe.log | tail -12
timeout 600 python -m pytest tests -m gpu -q -x -k "chol or sparsegpt or hessian or lora" 2>&1 | tail -5
