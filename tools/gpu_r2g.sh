#!/bin/bash
# bring-up build: epilogue cycle profile of one steady-state fp16 GEMM (h epilogue, K = 384 and 768)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for K in 384 768; do F2G_PAIR_DBG=16 timeout 120 python tools/one_gemm.py h $K 4 2>&1 | grep "epi prof" | tail -2 | tee -a gpurun_out/epi_prof.log; done
F2G_PAIR_DBG=16 timeout 120 python tools/one_gemm.py res 768 4 2>&1 | grep "epi prof" | tail -2 | tee -a gpurun_out/epi_prof.log
