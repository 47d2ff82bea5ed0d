"""Per-launch device timing (CUDA events, warm caches, eager) of one inference step at the bench shape."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from flow2gan_b200 import _lib as L
m = bench.build_model(torch.device("cuda", 0))
mel, noise = bench.synth_inputs()
rec = []
def wrap(name):
    fn = getattr(L, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        extra = ""
        if name == "gemm_group":
            d = a[0]; extra = " ".join(f"{x.M}x{x.N}x{x.K}/bn{x.bn}" for x in d)
        rec.append((name, extra, e0, e1)); return r
    return w
with torch.no_grad():
    plan = m.plan(bench.B, bench.FRAMES, bench.T, False)
    for _ in range(3): plan.infer(mel.cuda(), noise.cuda(), None, 1, False, use_graph=False)
    names = ["gemm_group", "block_pre", "stft", "irfft_frames", "ola_combine", "biasnorm", "linear_small_group", "time_sinusoid", "im2col_cf"]
    import flow2gan_b200.engine as E
    for n in names: setattr(L, n, wrap(n))
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); plan.infer(mel.cuda(), noise.cuda(), None, 1, False, use_graph=False); t1.record()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, extra, e0, e1 in rec:
    key = name + (" " + extra if extra else "")
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values())
print(f"eager step wall (events) {t0.elapsed_time(t1)*1e3:.0f} us; sum of launches {tot:.0f} us")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v[1]:8.1f} us  n={v[0]:3d} avg={v[1]/v[0]:7.1f}  {k[:150]}")
