#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for d in 0 8; do SCAN_EPIS=h SCAN_KS=384,768 F2G_PAIR_DBG=$d timeout 300 python tools/pair_f16_scan.py 2>&1 | tee -a gpurun_out/pair_f16_scan3.log; done
F2G_PAIR_DBG=8 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_generator_gpu.py -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
F2G_PAIR_DBG=8 timeout 600 python bench.py --no-legs --no-cpu --no-train > gpurun_out/bench_dbg8.json 2> gpurun_out/bench_dbg8.err
python - <<'P'
import json
for l in open('gpurun_out/bench_dbg8.json'):
    if l.startswith('{'):
        d = json.loads(l); print('dbg8: ms/step %.4f gemm frac %.3f (%.1f us/launch)' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['us_per_launch_avg']))
P
