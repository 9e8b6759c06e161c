#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sgpt_overlap_probe.py > gpurun_out/r02ad_sgpt_overlap_probe.log 2>&1; tail -12 gpurun_out/r02ad_sgpt_overlap_probe.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
