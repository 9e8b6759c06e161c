#!/bin/bash
# N-GPU default bench (driver's launch line); N from $1
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02zz_bench_wanda_nm_${N}gpu.json 2> gpurun_out/r02zz_bench_${N}gpu.err
echo "rc=$?"; tail -40 gpurun_out/r02zz_bench_${N}gpu.err | cut -c1-400
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02zz_bench_wanda_nm_${N}gpu.json') if l.startswith('{')][-1]
print("N=$N headline", round(d["value"]*1e3,3), "ms e2e", d["e2e"], d["roofline"]["frac"], d["clocks"])
for m,v in d["methods"].items(): print(m, round(v["value"]*1e3,3), v["roofline"].get("spans_ms_per_step"))
for k,v in d.get("workloads",{}).items(): print(k, v.get("value") if isinstance(v,dict) else v)
PY
