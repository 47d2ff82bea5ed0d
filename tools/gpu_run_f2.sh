#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
F2G_PDL=0 timeout 120 python tools/pre_bench.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
tail -1 gpurun_out/ncu_launch_warm.log
