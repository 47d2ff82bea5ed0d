#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -c 4 -f -o gpurun_out/prof_pair python tools/pair_prof.py > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
