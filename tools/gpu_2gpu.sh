#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_nccl_gpu.py tests/test_zz_properties_gpu.py::test_parity_on_trained_weight_proxy -q --tb=short -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|FAILED|rank|proxy|after|Error|assert" | cut -c1-250
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'P'
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d = json.loads(l); t = d.get('gan_train') or {}
        print('N=2: value %.1fM e2e %.1fM ms/step %.4f train %.2f ms/pair (%s)' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], t.get('ms_per_pair', 0), t.get('value')))
P
tail -3 gpurun_out/bench_n2.err
