#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/t1.log
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -s 2>&1 | grep -E "grad rel-err|fm loss|passed|failed|Error" > gpurun_out/t2.log
timeout 120 python tools/gemm_bench.py > gpurun_out/gb.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1
F2G_BN1=128 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1_bn128.log 2>&1
for f in t1 t2 gb bench_n1 bench_n1_bn128; do echo "== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-420; done
