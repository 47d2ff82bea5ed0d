#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
(timeout 300 python tools/merge_probe.py; F2G_PAIR_FORCE_GENERIC=1 timeout 300 python tools/merge_probe.py) 2>&1 | grep -v "^f16\|^   f16\|^dbg" | tee gpurun_out/merge_probe.log
