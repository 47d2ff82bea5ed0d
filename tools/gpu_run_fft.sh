#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "fft or stft" 2>&1 | tail -15 | cut -c1-300 > gpurun_out/t_fft.log
cat gpurun_out/t_fft.log
grep -q "failed\|error" gpurun_out/t_fft.log && exit 1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
for f in 1 0; do
F2G_FFT_SMEM=$f timeout 900 python bench.py > gpurun_out/bench_fft$f.json 2> gpurun_out/bench_fft$f.err
tail -3 gpurun_out/bench_fft$f.err
python - $f <<'P'
import json, sys
for l in open('gpurun_out/bench_fft%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('smem-fft', sys.argv[1], 'ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f other %.0f TF/s train %.1f ms' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], d['roofline_other']['achieved'], d['gan_train']['ms_per_pair']))
P
done
cp gpurun_out/bench_fft0.json gpurun_out/bench_full.json
