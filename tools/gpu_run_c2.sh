#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_trainer_gpu.py -x -q -m gpu -s 2>&1 | tail -30 | cut -c1-400
echo "== train pair (graph)"; timeout 300 python tools/one_train_pair.py 2>&1 | tail -3
