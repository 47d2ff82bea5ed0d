#!/bin/bash
# 2- and 4-step inference bench lines (BASELINE.json configs[1])
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for n in 2 4; do
timeout 600 python bench.py --no-train --n-timesteps $n --steps 30 > gpurun_out/bench_steps$n.json 2> gpurun_out/bench_steps$n.err
tail -2 gpurun_out/bench_steps$n.err
python - $n <<'P'
import json, sys
for l in open('gpurun_out/bench_steps%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print(sys.argv[1], 'steps: ms/call %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f cpu %.2fM' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], d['cpu_baseline']['value']/1e6))
P
done
