#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python scripts/full_model_dp_probe.py wanda_nm > gpurun_out/r02y_full_model_probe_1gpu.log 2>&1; grep -E "^rep|cumulative|prune|hook|add_batch|merge|all_reduce|capture|stack|synchronize|\.py" gpurun_out/r02y_full_model_probe_1gpu.log | head -60
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/full_model_dp_probe.py wanda_nm > gpurun_out/r02y_full_model_probe_${N}gpu.log 2>&1; grep -E "^rep|cumulative|prune|hook|add_batch|merge|all_reduce|capture|stack|synchronize|\.py" gpurun_out/r02y_full_model_probe_${N}gpu.log | head -60
