#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python tools/train_gemm_census.py 2>&1 | tail -60 | tee gpurun_out/train_gemm_census.log
