#!/bin/bash
# Round-2 first hardware pass (under gpurun, one B200): whole -m gpu suite WITHOUT -x (every failure
# listed), default bench line, the opt-in switches against it, the L2->SM fabric probe.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all_full.log 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/t_all_full.log | cut -c1-250 | tee gpurun_out/t_all.log
line() { python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        h = d.get('roofline_hbm') or {}
        print('%s: ms/step %.3f e2e %.1fM gemm frac %.3f hbm frac %s train %s' % (
            sys.argv[1], d['ms_per_step'], d['e2e']['value'] / 1e6, d['roofline']['frac'],
            ('%.3f' % h['frac']) if 'frac' in h else h.get('error'), d.get('gan_train', {})))
P
}
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; line gpurun_out/bench_default.json
F2G_FUSED_LOSSES=1 timeout 900 python bench.py > gpurun_out/bench_fused_losses.json 2> gpurun_out/bench_fused_losses.err; line gpurun_out/bench_fused_losses.json
F2G_PAIR_BN_HINT=1 timeout 600 python bench.py --no-train > gpurun_out/bench_bn_hint.json 2> gpurun_out/bench_bn_hint.err; line gpurun_out/bench_bn_hint.json
F2G_F16_COND=1 timeout 600 python bench.py --no-train > gpurun_out/bench_f16_cond.json 2> gpurun_out/bench_f16_cond.err; line gpurun_out/bench_f16_cond.json
F2G_BLOCK_OPERANDS=tf32 timeout 600 python bench.py --no-train > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; line gpurun_out/bench_tf32.json
for c in 0 1; do
  F2G_CACHE_TIME=$c timeout 600 python bench.py --no-train --n-timesteps 4 --steps 30 > gpurun_out/bench_n4_cache$c.json 2> gpurun_out/bench_n4_cache$c.err; line gpurun_out/bench_n4_cache$c.json
done
bash tools/gpu_run_fabric.sh
