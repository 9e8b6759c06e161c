#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --method wanda_unstructured --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02final5_bench_wanda_unstructured_4gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r02final5_bench_wanda_unstructured_4gpu.json').readline()); print('4 GPUs', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'], 'frac', d['roofline']['frac'])"
