#!/bin/bash
# A/B: hardware queue count for the concurrent SparseGPT chains (14+ streams vs the default 8 connections)
mkdir -p gpurun_out
for mc in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 600 python bench.py --method sparsegpt --no-other-methods --no-cpu-baseline --no-full-model --steps 4 --warmup 3 > gpurun_out/r02m_sparsegpt_mc$mc.json 2> gpurun_out/r02m_sparsegpt_mc$mc.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02m_sparsegpt_mc$mc.json'))
print("max_connections $mc: sparsegpt", round(d["value"]*1e3,2), "ms/block", {k: round(v,2) for k,v in d["roofline"]["spans_ms_per_step"].items()}, d["clocks"], "e2e", d["e2e"]["value"])
PY
done
