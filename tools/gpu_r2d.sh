#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ttp tools/tma_tensor_probe.cu -lcuda && timeout 300 /tmp/ttp | tee gpurun_out/tma_tensor_probe.log
for d in 0 1 3; do F2G_PAIR_DBG=$d timeout 300 python tools/pair_f16_scan.py 2>&1 | tee -a gpurun_out/pair_f16_scan.log; done
timeout 600 python -m pytest tests/test_chain_guard_gpu.py tests/test_trainer_gpu.py -q --tb=short -p no:cacheprovider 2>&1 | tail -15
