#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dsnot" 2>&1 | tail -3
rm -f gpurun_out/r02ah_dsnot.log
for C in 4096 11008; do timeout 300 python scripts/dsnot_ncu.py $C 2>&1 | tail -2 >> gpurun_out/r02ah_dsnot.log; done; cat gpurun_out/r02ah_dsnot.log
timeout 600 python bench.py --method dsnot --no-other-methods --no-cpu-baseline --no-full-model --steps 5 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('dsnot', round(d['value']*1e3,3), 'ms/block', d['roofline']['spans_ms_per_step'])"
