#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for v in 0 1 2 3; do F2G_PRE_VARIANT=$v timeout 120 python tools/pre_bench.py 2>&1 | tail -2; done
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
