"""One eager (non-graph) 1-step inference at the bench shape, after a warm-up, for ncu launch lists:
   ncu --metrics gpu__time_duration.sum ... python tools/one_step.py
Only the launches between the two cudaProfilerStart/Stop markers are the measured step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
m = bench.build_model(torch.device("cuda", 0))
mel, noise = bench.synth_inputs()
with torch.no_grad():
    plan = m.plan(bench.B, bench.FRAMES, bench.T, False)
    for _ in range(2):
        plan.infer(mel.cuda(), noise.cuda(), None, 1, False, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.infer(mel.cuda(), noise.cuda(), None, 1, False, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
