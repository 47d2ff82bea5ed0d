#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gan_gpu.py -q -m gpu -s -k "conv2d" 2>&1 | grep -E "conv2d_cl|passed|failed" | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
tail -1 gpurun_out/ncu_launch_warm.log
