#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsnot_walk2 -c 1 -s 1 -o gpurun_out/r02s_dsnot_walk2 python scripts/dsnot_ncu.py > gpurun_out/r02s_ncu.log 2>&1; tail -2 gpurun_out/r02s_ncu.log
timeout 600 python bench.py --method dsnot --no-other-methods --no-cpu-baseline --no-full-model --steps 5 --warmup 3 > gpurun_out/r02s_bench_dsnot.json 2> gpurun_out/r02s_bench_dsnot.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02s_bench_dsnot.json'))
print("dsnot", round(d["value"]*1e3,3), "ms/block", d["roofline"]["spans_ms_per_step"], d["clocks"])
PY
