#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 3 -c 1 -f -o gpurun_out/prof_gemm_h384 python tools/one_gemm.py h 384 4 > gpurun_out/ncu_g1.log 2>&1; tail -2 gpurun_out/ncu_g1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 3 -c 1 -f -o gpurun_out/prof_gemm_res768 python tools/one_gemm.py res 768 4 > gpurun_out/ncu_g2.log 2>&1; tail -2 gpurun_out/ncu_g2.log
