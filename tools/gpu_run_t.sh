#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for v in "F2G_PAIR_DBG=0" "F2G_PAIR_DBG=1" "F2G_PAIR_DBG=2" "F2G_GEMM_V1=1"; do
  env $v timeout 120 python tools/pair_bench.py 2>&1 | tail -14
done | tee gpurun_out/pair_bench.log
