#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:block_pre --launch-skip 4 -c 2 -f -o gpurun_out/prof_pre python tools/one_step.py > gpurun_out/ncu_pre.log 2>&1
tail -2 gpurun_out/ncu_pre.log
