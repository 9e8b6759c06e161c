#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stats or sqnorm or wanda or batch" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-full-model --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('headline', round(d['value']*1e3,4), 'ms/block eager', round(d['config']['eager_ms_per_step'],4), d['roofline']['spans_ms_per_step'], 'frac', d['roofline']['frac'])
for m,v in d['methods'].items(): print(' ', m, round(v['value']*1e3,3), v['roofline'].get('spans_ms_per_step'))"
