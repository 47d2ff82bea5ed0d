#!/bin/bash
# windowed-conv bring-up: conv tests first, then the whole GPU suite + the default bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gan_gpu.py -q -x -s -k "conv2d" 2>&1 | grep -E "conv2d|windowed|passed|failed|FAILED|Error|error|assert" | cut -c1-400 > gpurun_out/t_win.log
cat gpurun_out/t_win.log | tail -40
grep -q failed gpurun_out/t_win.log && exit 1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python - <<'P'
import json
for l in open('gpurun_out/bench_full.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac']))
        print('train', d.get('gan_train')); print('cpu', d.get('cpu_baseline'))
P
