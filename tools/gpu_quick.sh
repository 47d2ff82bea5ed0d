#!/bin/bash
# quick hardware check of a kernel change: kernel + generator + GAN parity tests, headline bench without the extra legs
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py tests/test_gan_gpu.py tests/test_train_gpu.py -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --no-legs --no-cpu ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python tools/print_quick.py
tail -3 gpurun_out/bench_quick.err
