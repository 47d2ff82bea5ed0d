#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --durations=15 > gpurun_out/t_all_full.log 2>&1
grep -E "passed|failed|FAILED|ERROR|^E  |s call|s setup" gpurun_out/t_all_full.log | cut -c1-260 | head -80
