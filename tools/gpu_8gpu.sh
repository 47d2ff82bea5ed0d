#!/bin/bash
# N=8 scaling line of bench.py under torchrun (the driver's round-end launch), bounded.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - $N <<'P'
import json, sys
n = sys.argv[1]
for l in open(f'gpurun_out/bench_n{n}.json'):
    if l.startswith('{'):
        d = json.loads(l); t = d.get('gan_train') or {}
        print('N=%s: value %.1fM e2e %.1fM ms/step %.4f train %.2f ms/pair (%s)' % (n, d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], t.get('ms_per_pair', 0), t.get('value')))
P
tail -3 gpurun_out/bench_n$N.err
