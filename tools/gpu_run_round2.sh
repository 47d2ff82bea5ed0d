#!/bin/bash
# Round-2 kick-off (under gpurun, one B200): everything that was written after round 1's GPU budget
# was spent gets its first hardware run, in the order of what later steps depend on.
#   1. full -m gpu suite (the test_zz_* files are the new ones)
#   2. default bench line (adds the roofline_hbm leg)
#   3. opt-in switches, each against the default: fused GAN loss kernels, chained-GEMM N-tile hint
#   4. L2 -> SM delivery probe (unicast vs multicast, cluster sizes 1-8)
#   5. torch-glue census of the training pair
#   6. ncu profile set (launch lists + one --set full capture of the GEMM launches)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error|error" | cut -c1-300 | tee gpurun_out/t_all.log
line() { python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        h = d.get('roofline_hbm') or {}
        print('%s: ms/step %.3f e2e %.1fM gemm frac %.3f hbm frac %s train %.1f ms' % (
            sys.argv[1], d['ms_per_step'], d['e2e']['value'] / 1e6, d['roofline']['frac'],
            ('%.3f' % h['frac']) if 'frac' in h else h.get('error'), d.get('gan_train', {}).get('ms_per_pair', float('nan'))))
P
}
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; line gpurun_out/bench_default.json
F2G_FUSED_LOSSES=1 timeout 900 python bench.py > gpurun_out/bench_fused_losses.json 2> gpurun_out/bench_fused_losses.err; line gpurun_out/bench_fused_losses.json
F2G_PAIR_BN_HINT=1 timeout 900 python bench.py --no-train > gpurun_out/bench_bn_hint.json 2> gpurun_out/bench_bn_hint.err
python - <<'P'
import json
for l in open('gpurun_out/bench_bn_hint.json'):
    if l.startswith('{'):
        d = json.loads(l); print('bn hint: ms/step %.3f gemm frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
P
F2G_F16_COND=1 timeout 900 python bench.py --no-train > gpurun_out/bench_f16_cond.json 2> gpurun_out/bench_f16_cond.err
python - <<'P'
import json
for l in open('gpurun_out/bench_f16_cond.json'):
    if l.startswith('{'):
        d = json.loads(l); print('fp16 cond rows: ms/step %.3f e2e %.1fM' % (d['ms_per_step'], d['e2e']['value'] / 1e6))
P
for n in 2 4; do
  for c in 0 1; do
    F2G_CACHE_TIME=$c timeout 900 python bench.py --no-train --n-timesteps $n --steps 30 > gpurun_out/bench_n${n}_cache$c.json 2> gpurun_out/bench_n${n}_cache$c.err
    python - gpurun_out/bench_n${n}_cache$c.json <<'P'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'])
P
  done
done
bash tools/gpu_run_fabric.sh
timeout 900 python tools/train_glue_census.py 30 > gpurun_out/train_glue_census.log 2>&1; head -20 gpurun_out/train_glue_census.log
[ "$1" = "noprof" ] && exit 0
bash tools/gpu_profile.sh
