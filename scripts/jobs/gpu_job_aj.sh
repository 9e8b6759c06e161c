#!/bin/bash
mkdir -p gpurun_out
for b in 0 1 0 1; do
VLMC_BENCH_STATS_BATCH=$b timeout 600 python bench.py --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('stats_batch=$b', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), 'graph', d['config'].get('graph_ms_per_step'), d['roofline']['spans_ms_per_step'])"
done
