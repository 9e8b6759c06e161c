#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stats or dsnot or batch_statistics or shared_inputs" 2>&1 | tail -5
for b in 0 1; do
VLMC_BENCH_STATS_BATCH=$b timeout 600 python bench.py --method dsnot_elided --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('dsnot_elided stats_batch=$b', round(d['value']*1e3,4), 'ms/block', d['roofline']['spans_ms_per_step'])"
done
