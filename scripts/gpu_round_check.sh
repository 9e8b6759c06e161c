#!/bin/bash
# What the driver runs at round end, in one call: GPU parity suite, smoke(), reference arm, default bench line.
TAG=${1:-r02final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; head -c 400 gpurun_out/${TAG}_bench_reference_arm.json; echo
timeout 1500 python bench.py > gpurun_out/${TAG}_bench_wanda_nm_1gpu.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/${TAG}_bench_wanda_nm_1gpu.json') if l.startswith('{')][-1]
print("headline", round(d["value"]*1e3,3), "ms  e2e", d["e2e"]["value"], "frac", round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"], d["clocks"])
for m,v in d["methods"].items(): print(" ", m, round(v["value"]*1e3,3), v["roofline"].get("spans_ms_per_step"), "traffic", v["roofline"].get("traffic"))
for k,v in d.get("workloads",{}).items(): print(" ", k, (v.get("value"), v.get("error")) if isinstance(v,dict) else v)
print(" cpu_baseline", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("kind"))
PY
