#!/bin/bash
mkdir -p gpurun_out
for w in 1 8 32; do
VLMC_STATS_BATCH_WAVES=$w timeout 600 python bench.py --method wanda_nm_shared --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('shared waves=$w', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'])"
done
