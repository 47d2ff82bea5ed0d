#!/bin/bash
# 2-GPU bench line (replica inference + data-parallel GAN train pair with NCCL gradient all-reduce)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err
python - $N <<'P'
import json, sys
for l in open('gpurun_out/bench_n%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms/step %.3f value %.1fM e2e %.1fM' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
        print('train', d.get('gan_train'))
P
