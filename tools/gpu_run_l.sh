#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-300 > gpurun_out/t_all.log
timeout 120 python tools/gemm_bench.py > gpurun_out/gb.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1
F2G_BN1=128 timeout 600 python bench.py --steps 50 --warmup 5 --no-train > gpurun_out/bench_n1_bn128.log 2>&1
for f in t_all gb bench_n1 bench_n1_bn128; do echo "== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-2600; done
