#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "rowselect or wanda or toy or dsnot" 2>&1 | tail -4
for b in 1 0 1 0; do
VLMC_BENCH_SELECT_BATCH=$b timeout 300 python bench.py --method wanda_unstructured --no-other-methods --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('select batch=$b', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'])"
done
