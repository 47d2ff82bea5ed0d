#!/bin/bash
# Round-2 second hardware pass: whole -m gpu suite (no -x), both bench arms, fabric probe v2,
# first --set full capture of the step's HBM-bound kernels, torch-glue census of the train pair.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all_full.log 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/t_all_full.log | cut -c1-250 | tee gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
bash tools/gpu_run_fabric.sh
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"block_pre|stft_group_warp|irfft_group_warp|ola_combine|biasnorm" -c 8 -f -o gpurun_out/prof_hbm_step python tools/one_step.py > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
timeout 900 python tools/train_glue_census.py 40 > gpurun_out/train_glue_census.log 2>&1; head -45 gpurun_out/train_glue_census.log
