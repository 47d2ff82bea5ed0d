#!/bin/bash
# full GPU suite + bench line, then (if green) the round's inference profile set: launch lists
# (cold = ncu default cache flush, and warm) and one --set full capture of the step's GEMM launches
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
grep -q "failed\|Error" gpurun_out/t_all.log && exit 1
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python - <<'P'
import json
for l in open('gpurun_out/bench_full.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms/step %.3f value %.1fM e2e %.1fM gemm %.0f TF/s frac %.3f other %.0f TF/s train %.1f ms' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['achieved'], d['roofline']['frac'], (d.get('roofline_other') or {'achieved': 0})['achieved'], d['gan_train']['ms_per_pair']))
P
[ "$1" = "noprof" ] && exit 0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_warm.csv python tools/one_step.py > gpurun_out/ncu_launch_warm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ -c 40 -f -o gpurun_out/prof_gemm_step python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
