#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_warm.csv python tools/one_train_pair.py > gpurun_out/ncu_train_warm.log 2>&1
tail -1 gpurun_out/ncu_train_warm.log
