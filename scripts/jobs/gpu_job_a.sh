#!/bin/bash
# r02a: GPU parity suite (all failures collected), chain timings, reference arm + default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc > gpurun_out/r02a_nproc.txt
timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_nothing 2>&1 | tail -40 > gpurun_out/r02a_pytest_gpu.log; tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 300 python scripts/chol_probe.py > gpurun_out/r02a_chol_probe.log 2>&1; tail -12 gpurun_out/r02a_chol_probe.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err; tail -c 600 gpurun_out/r02a_bench_reference.json
