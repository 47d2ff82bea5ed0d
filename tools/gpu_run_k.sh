#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|whole-vector" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
bash tools/gpu_profile.sh
