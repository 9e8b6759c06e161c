#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mask_pack or rowselect or packed" 2>&1 | tail -4
for b in 1 0; do
VLMC_BENCH_SELECT_BATCH=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --method wanda_unstructured --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('2 GPUs select batch=$b', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'])"
done
