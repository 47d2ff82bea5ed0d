#!/bin/bash
# GEMM-focused check: kernel tests first (fail fast), then the whole GPU suite, then the bench.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-400 > gpurun_out/t_kern.log
cat gpurun_out/t_kern.log
grep -q failed gpurun_out/t_kern.log && exit 1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import sys, json
for l in open('gpurun_out/bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print('  ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f train %s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], (d.get('gan_train') or {}).get('ms_per_pair')))
"
