#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/k.log
timeout 900 python -m pytest tests/test_generator_gpu.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/g.log
timeout 120 python tools/gemm_bench.py > gpurun_out/gb.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1
F2G_BN1=128 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_bn128.log 2>&1
F2G_BN2=64 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_bn2_64.log 2>&1
F2G_BN2=256 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_bn2_256.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
for f in k g gb bench_n1 bench_n1_bn128 bench_n1_bn2_64 bench_n1_bn2_256; do echo "== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-900; done
