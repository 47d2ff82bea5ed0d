#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
for cfg in "256 128" "128 128" "128 64" "256 64"; do set -- $cfg; echo "BN1=$1 BN2=$2"; F2G_BN1=$1 F2G_BN2=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-train 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved']))
"; done
