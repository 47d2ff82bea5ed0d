#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gan_gpu.py -q -m gpu -s --tb=line 2>&1 | grep -E "conv2d_cl|losses|grad rel-err|worst|passed|failed|Error|error" | cut -c1-1800 > gpurun_out/t3.log
cat gpurun_out/t3.log
