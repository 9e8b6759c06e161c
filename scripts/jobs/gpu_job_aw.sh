#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-full-model > gpurun_out/r02final3_bench_wanda_nm_8gpu.json 2> gpurun_out/r02final3_bench_8gpu.err
tail -2 gpurun_out/r02final3_bench_8gpu.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02final3_bench_wanda_nm_8gpu.json') if l.startswith('{')][-1]
print("headline", round(d["value"]*1e3,3), "ms  e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"])
for m,v in d["methods"].items(): print(" ", m, round(v["value"]*1e3,3), v["roofline"].get("spans_ms_per_step"))
PY
