#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02v_pytest_gpu.log; cat gpurun_out/r02v_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r02v_bench_wanda_nm_1gpu.json 2> gpurun_out/r02v_bench.err; tail -3 gpurun_out/r02v_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02v_bench_wanda_nm_1gpu.json'))
print("headline", d["value"]*1e3, "ms e2e", d["e2e"]["value"], d["roofline"]["frac"])
for m,v in d["methods"].items(): print(m, round(v["value"]*1e3,3), v["roofline"].get("spans_ms_per_step"))
for k,v in d.get("workloads",{}).items(): print(k, v.get("value"))
PY
