#!/bin/bash
mkdir -p gpurun_out
for w in 16 24 32 48 64; do
VLMC_STATS_BATCH_WAVES=$w timeout 600 python bench.py --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('waves=$w', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'])"
done
