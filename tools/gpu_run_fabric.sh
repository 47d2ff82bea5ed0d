#!/bin/bash
# (under gpurun) L2 -> SM delivery probe: distinct / shared-unicast / shared-multicast loads at
# cluster sizes 1, 2, 4, 8 -- decides whether the GEMM's operand-byte bound moves with multicast
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fabric_probe tools/fabric_probe.cu 2> gpurun_out/fabric_probe.build.log || { tail -5 gpurun_out/fabric_probe.build.log; exit 1; }
timeout 120 /tmp/fabric_probe | tee gpurun_out/fabric_probe.log
