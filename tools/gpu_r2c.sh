#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/trp tools/tma_rate_probe.cu && timeout 300 /tmp/trp | tee gpurun_out/tma_rate_probe.log
timeout 900 python tools/arm_matrix.py --oracle 2>&1 | tee gpurun_out/arm_matrix.log
