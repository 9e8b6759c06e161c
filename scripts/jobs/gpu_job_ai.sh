#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stats or wanda or sqnorm" 2>&1 | tail -3
timeout 300 python scripts/sqnorm_probe.py 2>&1 | tee gpurun_out/r02ai_sqnorm_probe.log
