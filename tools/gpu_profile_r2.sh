#!/bin/bash
# Round-2 evidence capture (under gpurun, one B200).  Post-process here with
#   python tools/summarize_profiles.py r02 ; python tools/summarize_hbm.py r02 gpurun_out/prof_hbm_*.ncu-rep
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N="ncu --clock-control none --profile-from-start off"
# (1) launch lists: one inference step (cold = ncu default, and warm caches), one GAN D+G pair
timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
timeout 900 $N --metrics gpu__time_duration.sum --cache-control none --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
timeout 1200 $N --metrics gpu__time_duration.sum --cache-control none --csv --log-file gpurun_out/launches_train.csv python tools/one_train_pair.py > gpurun_out/ncu_train.log 2>&1
# (2) --set full: the step's tcgen05 GEMM launches (dominant kernel), its HBM-bound kernels
timeout 900 $N --set full --import-source on -k regex:gemm_ -c 20 -f -o gpurun_out/prof_gemm_step python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
timeout 900 $N --set full -k regex:"block_pre|stft_group_warp|irfft_group_warp|ola_combine|biasnorm|linear_small|im2col_cf" -c 24 -f -o gpurun_out/prof_hbm_step python tools/one_step.py > gpurun_out/ncu_hbm1.log 2>&1
# (3) --set full: HBM-bound kernels of the train pair, two launches each
for k in act_bwd_win act_bwd_vec adam_update adam_reduce block_bwd_c block_bwd_a pad2d conv_small_fwd conv_small_wgrad_kernel conv_small_dgrad spec_loss_bwd stft_kernel loss_terms_fwd loss_terms_bwd; do
  timeout 600 $N --set full -k regex:$k -c 2 -f -o gpurun_out/prof_hbm_train_$k python tools/one_train_pair.py > gpurun_out/ncu_hbm_$k.log 2>&1
done
# (4) data-path / model-average kernels
timeout 600 $N --set full -k regex:"average_update|gain_resample|pcm16_encode" -c 4 -f -o gpurun_out/prof_hbm_datapath python tools/one_datapath.py > gpurun_out/ncu_hbm_dp.log 2>&1
# (5) summarise ON THE BOX (only gpurun_out/ comes back, 64 MiB at most): tables + JSON into gpurun_out/profiles_out/,
#     then drop the raw reports except the GEMM capture (source-level view of the dominant kernel)
export F2G_PROFILES_OUT=gpurun_out/profiles_out
python tools/summarize_profiles.py r02 > /dev/null
python tools/summarize_hbm.py r02 gpurun_out/prof_hbm_step.ncu-rep gpurun_out/prof_hbm_train_*.ncu-rep gpurun_out/prof_hbm_datapath.ncu-rep > /dev/null
ncu -i gpurun_out/prof_gemm_step.ncu-rep --page source --csv > gpurun_out/profiles_out/r02_gemm_step_source.csv 2>/dev/null
rm -f gpurun_out/prof_hbm_*.ncu-rep
for f in gpurun_out/ncu_full.log gpurun_out/ncu_hbm1.log gpurun_out/ncu_hbm_dp.log; do tail -n 2 $f; done
ls -la gpurun_out/*.ncu-rep gpurun_out/profiles_out | awk '{print $5, $9}'
