#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
for f in 0 1; do
F2G_PAIR_FILL=$f timeout 900 python bench.py > gpurun_out/bench_fill$f.json 2> gpurun_out/bench_fill$f.err
tail -3 gpurun_out/bench_fill$f.err
python - $f <<'P'
import json, sys
for l in open('gpurun_out/bench_fill%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('fill', sys.argv[1], 'ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f other %.0f TF/s train %.1f ms' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], d['roofline_other']['achieved'], d['gan_train']['ms_per_pair']))
P
done
cp gpurun_out/bench_fill1.json gpurun_out/bench_full.json
