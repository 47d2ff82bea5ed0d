#!/bin/bash
# Round profile capture (run under gpurun): launch list of one inference step (cold cache = ncu
# default, and warm), one full capture of the dominant kernel (the step's GEMM launches), the
# GAN train-pair launch list.  Post-process with tools/summarize_profiles.py <tag>.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ -c 40 -f -o gpurun_out/prof_gemm_step python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python tools/one_train_pair.py > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log gpurun_out/ncu_train.log
