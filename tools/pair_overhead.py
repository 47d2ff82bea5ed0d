"""Fixed per-launch overhead and per-tile cost of the CTA-pair GEMM: one-tile problems, then
1/2/3/4 full waves of identical 256x256xK tiles (intercept = launch overhead + pipeline
fill/drain, slope = steady-state tile time: max(MMA, epilogue)).   python tools/pair_overhead.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()
dev = "cuda"


def timeit(descs, reps=50, graph=True):
    for _ in range(3): L.gemm_group(descs)
    torch.cuda.synchronize()
    if graph:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps): L.gemm_group(descs)
            g.replay(); s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); s.synchronize()
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): L.gemm_group(descs)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def prob(M, N, K, epi):
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev) * 0.02; c = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev); sl = torch.rand(N, device=dev)
    kw = {}
    if epi == "act":
        kw = dict(bias=bias.data_ptr(), slope=sl.data_ptr(), act=L.ACT_PRELU, round_tf32=1)
    elif epi == "res":
        kw = dict(bias=bias.data_ptr(), res=c.data_ptr(), ld_res=N, res_scale=sl.data_ptr())
    return L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, K, N, **kw), (a, b, c, bias, sl)


def main():
    for epi in ("plain", "act", "res"):
        d, keep = prob(256, 256, 32, epi)
        print(f"{epi:5s} one tile K=32      {timeit([d]):7.2f} us (graph)  {timeit([d], graph=False):7.2f} us (stream)", flush=True)
        for K in (384, 768, 2304):
            row = []
            for waves in (1, 2, 3, 4, 8):
                d, keep = prob(256 * 74 * waves // 4, 1024, K, epi)      # 74*waves tiles of 256x256
                us = timeit([d], reps=20)
                row.append(us)
            slope = (row[4] - row[3]) / 4
            print(f"{epi:5s} K={K:4d} waves 1,2,3,4,8: " + " ".join(f"{u:7.2f}" for u in row) +
                  f"   slope {slope:6.2f} us/wave  intercept {row[0] - slope:6.2f} us   "
                  f"mma-ideal {2*256*256*K/11.34e6:5.2f} us/tile", flush=True)


if __name__ == "__main__":
    main()
