#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/t_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
for f in t_all bench_n1 smoke; do echo "== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-3000; done
