#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "chained or f16" 2>&1 | tail -15 | cut -c1-300 > gpurun_out/t_chain.log
cat gpurun_out/t_chain.log
grep -q "failed\|error" gpurun_out/t_chain.log && exit 1
timeout 600 python -m pytest tests/test_generator_gpu.py -q -x -s 2>&1 | grep -E "rel|passed|failed|Error|assert" | cut -c1-300 | tail -20
timeout 900 python bench.py --no-train > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err
tail -3 gpurun_out/bench_chain.err
python - <<'P'
import json
for l in open('gpurun_out/bench_chain.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f launches/step %d' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches']/d['steps']))
P
