#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --tee 3 --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_full_${N}gpu.out 2> gpurun_out/r02x_full_${N}gpu.err
echo "rc=$?"; tail -30 gpurun_out/r02x_full_${N}gpu.err | cut -c1-300; echo ----; cut -c1-300 gpurun_out/r02x_full_${N}gpu.out | tail -20
