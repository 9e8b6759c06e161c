#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "obs_fused" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q -x -k "sparsegpt or obs or fasterprune" 2>&1 | tail -4
for f in 0 1; do echo "VLMC_OBS_FUSED=$f"; VLMC_OBS_FUSED=$f timeout 300 python scripts/chol_probe.py 2>&1 | grep obs_sweep; done | tee gpurun_out/r02ag_obs_fused_probe.log
for f in 0 1; do echo "VLMC_OBS_FUSED=$f"; VLMC_OBS_FUSED=$f timeout 300 python scripts/concurrency_probe.py 2>&1 | grep -E "^sweep"; done | tee -a gpurun_out/r02ag_obs_fused_probe.log
