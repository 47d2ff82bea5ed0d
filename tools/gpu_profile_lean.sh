#!/bin/bash
# Lean evidence capture (one B200, ~5 min): launch lists + the --set full GEMM capture, summarised on the box;
# only small files come back (the .ncu-rep is dropped, the source page is gzipped).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N="ncu --clock-control none --profile-from-start off"
timeout 600 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
timeout 600 $N --metrics gpu__time_duration.sum --cache-control none --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
timeout 900 $N --metrics gpu__time_duration.sum --cache-control none --csv --log-file gpurun_out/launches_train.csv python tools/one_train_pair.py > gpurun_out/ncu_train.log 2>&1
timeout 900 $N --set full --import-source on -k regex:gemm_ -c 20 -f -o gpurun_out/prof_gemm_step python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
export F2G_PROFILES_OUT=gpurun_out/profiles_out
python tools/summarize_profiles.py r02 > /dev/null
ncu -i gpurun_out/prof_gemm_step.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/profiles_out/r02_gemm_step_source.csv.gz
rm -f gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu_full.log; du -sh gpurun_out
