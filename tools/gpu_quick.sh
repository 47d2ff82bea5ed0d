#!/bin/bash
# quick hardware check of a kernel change: kernel + generator + GAN parity tests, headline bench without the extra legs
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py tests/test_gan_gpu.py tests/test_train_gpu.py -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --no-legs --no-cpu ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'P'
import json
for l in open('gpurun_out/bench_quick.json'):
    if l.startswith('{'):
        d = json.loads(l)
        t = d.get('gan_train') or {}
        print('ms/step %.4f e2e %.1fM gemm frac %.3f (%.1f us/launch) other %.3f hbm %.3f train %.2f ms/pair gemm_ms %s' % (
            d['ms_per_step'], d['e2e']['value'] / 1e6, d['roofline']['frac'], d['roofline']['us_per_launch_avg'],
            d.get('roofline_other', {}).get('frac', 0), d.get('roofline_hbm', {}).get('frac', 0),
            t.get('ms_per_pair', 0), (t.get('roofline') or {}).get('gemm_ms')))
P
tail -3 gpurun_out/bench_quick.err
