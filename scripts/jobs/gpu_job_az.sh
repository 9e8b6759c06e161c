#!/bin/bash
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o gpurun_out/r02ay_$1 -f python scripts/sgpt_chain_once.py > gpurun_out/r02ay_$1.log 2>&1
  tail -1 gpurun_out/r02ay_$1.log
}
cap trailing_k128 'gemm3x_kernel<.int.256, .bool.0, .bool.0' 20
cap merge_top 'gemm3x_kernel<.int.128, .bool.1, .bool.1' 83
cap obs_far_k512 'gemm3x_kernel<.int.256, .bool.1, .bool.0' 10
ls -la gpurun_out/r02ay_*.ncu-rep
