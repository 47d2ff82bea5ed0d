#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_warm.csv python tools/one_train_pair.py > gpurun_out/ncu_train_warm.log 2>&1
tail -2 gpurun_out/ncu_train_warm.log
