#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py -q -m gpu 2>&1 | tail -8 > gpurun_out/t1.log
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -s 2>&1 | grep -E "grad rel-err|worst tensors|fm loss|passed|failed|Error" | cut -c1-1500 > gpurun_out/t2.log
timeout 900 python -m pytest tests/test_gan_gpu.py -q -m gpu -s 2>&1 | tail -60 | cut -c1-1500 > gpurun_out/t3.log
timeout 120 python tools/gemm_bench.py 2>&1 | head -6 > gpurun_out/gb.log
for f in t1 t2 t3 gb; do echo "== $f"; tail -n 40 gpurun_out/$f.log; done
