#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python -m pytest tests/test_gan_gpu.py -q -m gpu -k conv2d_cl -s 2>&1 | grep -E "conv2d_cl|passed|failed" | cut -c1-200
echo "== v1"; F2G_GEMM_V1=1 timeout 300 python -m pytest tests/test_gan_gpu.py -q -m gpu -k conv2d_cl -s 2>&1 | grep -E "conv2d_cl|passed|failed" | cut -c1-200
echo "== train pair default"; timeout 300 python tools/one_train_pair.py 2>&1 | tail -3
echo "== train pair v1"; F2G_GEMM_V1=1 timeout 300 python tools/one_train_pair.py 2>&1 | tail -3
