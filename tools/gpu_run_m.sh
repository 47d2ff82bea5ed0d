#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_train_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-300 > gpurun_out/t_gan.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.log 2>&1
for f in t_gan bench_n1; do echo "== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-2600; done
