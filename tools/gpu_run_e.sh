#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for d in 0 4 8 16 32; do F2G_GEMM_DBG=$d timeout 120 python tools/gemm_bench.py 2>&1 | head -6; done > gpurun_out/gb_dbg2.log 2>&1
cat gpurun_out/gb_dbg2.log
