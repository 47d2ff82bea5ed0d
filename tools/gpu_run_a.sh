#!/bin/bash
# bring-up batch A: kernel tests, MN-major sweep, generator parity, bench, ncu captures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not mn_major" 2>&1 | tail -40 > gpurun_out/k1.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "mn_major" 2>&1 | tail -40 > gpurun_out/k2.log
timeout 600 python tools/mn_sweep.py > gpurun_out/mn_sweep.log 2>&1
timeout 900 python -m pytest tests/test_generator_gpu.py -q -m gpu -s 2>&1 | tail -60 > gpurun_out/g.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --n-timesteps 2 > gpurun_out/bench_n2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --n-timesteps 4 > gpurun_out/bench_n4.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 30 -c 4 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
for f in k1 k2 g smoke bench_n1 bench_n2 bench_n4 bench_ref; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
