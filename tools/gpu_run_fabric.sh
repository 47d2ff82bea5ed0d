#!/bin/bash
# (under gpurun) L2 -> SM delivery probe: distinct / shared-unicast / shared-multicast bulk loads at
# cluster sizes 1, 2, 4, 8, for several pipeline depths, plus the plain LDG path -- decides whether the
# GEMM's operand-byte bound is the SM's ingest port (then only shared-memory reuse helps)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
: > gpurun_out/fabric_probe.log
for cfg in "4 32" "6 32" "12 16" "3 64"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DPROBE_STAGES=$1 -DPROBE_CHUNK_KB=$2 -o /tmp/fabric_probe tools/fabric_probe.cu 2> gpurun_out/fabric_probe.build.log || { tail -5 gpurun_out/fabric_probe.build.log; exit 1; }
  timeout 120 /tmp/fabric_probe | tee -a gpurun_out/fabric_probe.log
done
