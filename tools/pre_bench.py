"""Micro-benchmark of the grouped block prologue (three branches, bench shape) + grouped STFT/iRFFT."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()
dev = "cuda"
B, Fm = 16, 94
Fs = [95, 189, 377]; Cs = [768, 512, 384]; fac = [1, 2, 4]
descs, keep = [], []
for F, C, f in zip(Fs, Cs, fac):
    R = B * F
    x = torch.randn(R, C, device=dev); a1 = torch.empty(R, C, device=dev)
    dw = torch.randn(7, C, device=dev); db = torch.randn(C, device=dev); nb = torch.randn(C, device=dev) * 0.1
    ls = torch.tensor(0.3, device=dev); cp = torch.randn(B * Fm + 1, 8 * C, device=dev); ts = torch.randn(1, 8 * C, device=dev)
    keep += [x, a1, dw, db, nb, ls, cp, ts]
    descs.append(L.block_pre_desc(x, B, F, C, C, dw, db, nb, ls, None, cp, 8 * C, Fm, f, B * Fm, ts, 0, a1, C))
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
print("variant", os.environ.get("F2G_PRE_VARIANT", "0"), "block_pre x3: %.1f us" % t(lambda: L.block_pre_group(descs)))
T = 24064
xa = torch.randn(B, T, device=dev)
probs, probs2 = [], []
for n, hop in ((1024, 256), (512, 128), (256, 64)):
    F = 1 + T // hop; ld = (n + 2 + 3) // 4 * 4
    pin = torch.empty(B * F, ld, device=dev); fr = torch.empty(B * F, n, device=dev)
    keep += [pin, fr]
    probs.append((xa, pin, n, hop, F, B * F, T, ld)); probs2.append((pin, fr, n, 0, 0, B * F, ld, n))
print("stft x3: %.1f us   irfft x3: %.1f us" % (t(lambda: L.stft_group(probs, B, T, 1)), t(lambda: L.irfft_group(probs2))))
