#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s train %s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d.get('gan_train')))
"
