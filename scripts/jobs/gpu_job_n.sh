#!/bin/bash
# potrf_block2_kernel: correctness (factorisation tests), A/B timing against v1, chain concurrency
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "chol or sparsegpt or obs or fasterprune" 2>&1 | tail -15 > gpurun_out/r02n_pytest_gpu.log; tail -5 gpurun_out/r02n_pytest_gpu.log
for v in 1 0; do
  echo "VLMC_POTRF_V1=$v" >> gpurun_out/r02n_chol_probe.log
  VLMC_POTRF_V1=$v timeout 300 python scripts/chol_probe.py >> gpurun_out/r02n_chol_probe.log 2>&1
done
cat gpurun_out/r02n_chol_probe.log
timeout 600 python scripts/concurrency_probe.py 2>&1 | head -12 > gpurun_out/r02n_concurrency_probe.log; cat gpurun_out/r02n_concurrency_probe.log
