#!/bin/bash
# r02k: parity suite, default bench line (all methods, workloads, cpu baseline), Hessian slab sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02k_pytest_gpu.log; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02k_pytest_gpu.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python scripts/hessian_slab_probe.py > gpurun_out/r02k_hessian_slab_probe.log 2>&1; cat gpurun_out/r02k_hessian_slab_probe.log
timeout 1800 python bench.py > gpurun_out/r02k_bench_wanda_nm_1gpu.json 2> gpurun_out/r02k_bench.err; tail -c 1500 gpurun_out/r02k_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench_wanda_nm_1gpu.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
for m,v in d.get("methods",{}).items(): print(m, round(v["value"]*1e3,3), "ms", v["roofline"]["frac"], {k: round(x,3) for k,x in v["roofline"]["spans_ms_per_step"].items()})
for k,v in d.get("workloads",{}).items(): print(k, json.dumps(v)[:400])
print("cpu", json.dumps(d.get("cpu_baseline"))[:600])
print("calib", d["config"].get("calib_batch_sweep"))
print("peaks", d.get("peaks"))
PY
