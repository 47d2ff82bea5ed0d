#!/bin/bash
# Round-end rehearsal on one B200: what the driver runs (GPU suite with -x, smoke, both bench arms), outputs kept
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; cut -c1-300 gpurun_out/bench_r02_reference.json
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 2500 gpurun_out/bench_r02_n1.json; tail -3 gpurun_out/bench_r02_n1.err
