#!/bin/bash
mkdir -p gpurun_out
for pr in 0 1; do
  VLMC_SCHED_PRIORITY=$pr timeout 600 python bench.py --method sparsegpt --no-other-methods --no-cpu-baseline --no-full-model --steps 4 --warmup 3 > gpurun_out/r02l_sparsegpt_prio$pr.json 2> gpurun_out/r02l_sparsegpt_prio$pr.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02l_sparsegpt_prio$pr.json'))
print("prio $pr: sparsegpt", round(d["value"]*1e3,2), "ms/block", {k: round(v,2) for k,v in d["roofline"]["spans_ms_per_step"].items()}, d["clocks"], "e2e", d["e2e"]["value"])
PY
done
timeout 600 python -m pytest tests -m gpu -q -k "sparsegpt or obs or chol" 2>&1 | tail -3
for m in wanda_nm wanda_unstructured dsnot sparsegpt; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic_$m.csv python bench.py --one-step --method $m > /dev/null 2>&1
  echo "traffic $m: $(wc -l < gpurun_out/traffic_$m.csv) lines"
done
python scripts/ncu_traffic.py gpurun_out/traffic_wanda_nm.csv gpurun_out/traffic_wanda_unstructured.csv gpurun_out/traffic_dsnot.csv gpurun_out/traffic_sparsegpt.csv | tail -30
cp profiles/ncu_traffic.json gpurun_out/r02l_ncu_traffic.json
