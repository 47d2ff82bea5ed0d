#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 4 -c 1 -o gpurun_out/prof_gemm2 python tools/gemm_one.py 1520 2304 768 256 > gpurun_out/ncu_g2.log 2>&1
tail -3 gpurun_out/ncu_g2.log
