#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for d in 0 1 2 3 4 7; do F2G_GEMM_DBG=$d timeout 120 python tools/gemm_bench.py; done > gpurun_out/gb_dbg.log 2>&1
F2G_DESC_GLOBAL=1 timeout 120 python tools/gemm_bench.py > gpurun_out/gb_descglobal.log 2>&1
F2G_TMA_L2PROMO=0 timeout 120 python tools/gemm_bench.py > gpurun_out/gb_promo0.log 2>&1
F2G_TMA_L2PROMO=2 timeout 120 python tools/gemm_bench.py > gpurun_out/gb_promo2.log 2>&1
timeout 300 python tools/cpu_threads.py > gpurun_out/cpu_threads.log 2>&1
cat gpurun_out/gb_dbg.log | grep -v "^$" | head -80
