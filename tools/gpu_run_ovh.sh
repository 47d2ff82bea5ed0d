#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python tools/pair_overhead.py 2>&1 | tail -30 | tee gpurun_out/pair_overhead.log
