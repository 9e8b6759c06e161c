#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02final5_bench_wanda_nm_1gpu.json 2> gpurun_out/r02final5_bench.err; tail -2 gpurun_out/r02final5_bench.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02final5_bench_wanda_nm_1gpu.json') if l.startswith('{')][-1]
r=d["roofline"]
print("headline", round(d["value"]*1e3,3), "ms  e2e", d["e2e"]["value"], "frac", round(r["frac"],3), "per-linear", round(r.get("frac_per_linear_bytes",0),3), "traffic/alg", r.get("traffic_over_algorithmic"), d["clocks"], "launches", d.get("gpu_launches"))
for m,v in d["methods"].items(): print(" ", m, round(v["value"]*1e3,3), round(v["roofline"]["frac"],3), v["roofline"].get("spans_ms_per_step"))
for k,v in d.get("workloads",{}).items(): print(" ", k, (v.get("value"), v.get("error")) if isinstance(v,dict) else v)
PY
