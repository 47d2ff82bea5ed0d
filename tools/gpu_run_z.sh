#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
for v in 1 0; do
F2G_PDL=$v timeout 600 python bench.py --steps 50 --warmup 5 --no-train > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import sys, json
for l in open('gpurun_out/bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print('  PDL=$v ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac']))
"
done
