"""Micro-benchmark of the grouped GEMM launches of one ConvNeXt block at the bench shape
(pwconv1 group, pwconv2 group) + a few single problems.  Env F2G_PAIR_DBG / F2G_GEMM_V1 select
pipeline experiments.   python tools/pair_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()
dev = "cuda"
R = [1520, 3024, 6032]; C = [768, 512, 384]

def timeit(descs, reps=20, flush=None):
    for _ in range(3): L.gemm_group(descs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): L.gemm_group(descs)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = sum(2.0 * d.M * d.N * d.K for d in descs)
    return us, fl / us / 1e6

def block_groups():
    g1, g2, keep = [], [], []
    for r, c in zip(R, C):
        a = torch.randn(r, c, device=dev); w1 = torch.randn(3 * c, c, device=dev) * 0.02
        h = torch.empty(r, 3 * c, device=dev); w2 = torch.randn(c, 3 * c, device=dev) * 0.02
        x = torch.randn(r, c, device=dev); b1 = torch.randn(3 * c, device=dev); sl = torch.rand(3 * c, device=dev)
        b2 = torch.randn(c, device=dev); rs = torch.rand(c, device=dev)
        keep += [a, w1, h, w2, x, b1, sl, b2, rs]
        g1.append(L.gemm_desc(a.data_ptr(), w1.data_ptr(), h.data_ptr(), r, 3 * c, c, c, c, 3 * c, bn=128,
                              bias=b1.data_ptr(), slope=sl.data_ptr(), act=L.ACT_PRELU, round_tf32=1))
        g2.append(L.gemm_desc(h.data_ptr(), w2.data_ptr(), x.data_ptr(), r, c, 3 * c, 3 * c, 3 * c, c, bn=128,
                              bias=b2.data_ptr(), res=x.data_ptr(), ld_res=c, res_scale=rs.data_ptr()))
    return g1, g2, keep

def block_groups_f16():
    g1, g2, keep = [], [], []
    for r, c in zip(R, C):
        a = torch.randn(r, c, device=dev).half(); w1 = (torch.randn(3 * c, c, device=dev) * 0.02).half()
        h = torch.empty(r, 3 * c, device=dev, dtype=torch.float16); w2 = (torch.randn(c, 3 * c, device=dev) * 0.02).half()
        x = torch.randn(r, c, device=dev); b1 = torch.randn(3 * c, device=dev); sl = torch.rand(3 * c, device=dev)
        b2 = torch.randn(c, device=dev); rs = torch.rand(c, device=dev)
        keep += [a, w1, h, w2, x, b1, sl, b2, rs]
        g1.append(L.gemm_desc(a.data_ptr(), w1.data_ptr(), h.data_ptr(), r, 3 * c, c, c, c, 3 * c,
                              bias=b1.data_ptr(), slope=sl.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1))
        g2.append(L.gemm_desc(h.data_ptr(), w2.data_ptr(), x.data_ptr(), r, c, 3 * c, 3 * c, 3 * c, c,
                              bias=b2.data_ptr(), res=x.data_ptr(), ld_res=c, res_scale=rs.data_ptr(), ab_f16=1))
    return g1, g2, keep


def main():
    if os.environ.get("F2G_BENCH_F16", "1") == "1":
        g1, g2, keep = block_groups_f16()
        for name, g in (("f16 pwconv1 x3", g1), ("f16 pwconv2 x3", g2)):
            us, tf = timeit(g)
            print(f"{name:16s} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
            for i, d in enumerate(g):
                us, tf = timeit([d])
                print(f"   f16 branch {i} M={d.M} N={d.N} K={d.K} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
        for (M, N, K) in ((18944, 2304, 768), (18944, 768, 2304), (8192, 8192, 2048)):
            a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half(); c = torch.empty(M, N, device=dev)
            us, tf = timeit([L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, K, N, ab_f16=1)], reps=10)
            print(f"f16 plain M={M} N={N} K={K} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
    tag = "dbg=%s v1=%s" % (os.environ.get("F2G_PAIR_DBG", "0"), os.environ.get("F2G_GEMM_V1", "0"))
    g1, g2, keep = block_groups()
    for name, g in (("pwconv1 x3", g1), ("pwconv2 x3", g2)):
        us, tf = timeit(g)
        print(f"{tag} {name:12s} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
        for i, d in enumerate(g):
            us, tf = timeit([d])
            print(f"{tag}   branch {i} M={d.M} N={d.N} K={d.K} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
    for (M, N, K) in ((18944, 2304, 768), (18944, 768, 2304), (8192, 8192, 2048)):
        a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); c = torch.empty(M, N, device=dev)
        us, tf = timeit([L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, K, N, bn=256)], reps=10)
        print(f"{tag} plain M={M} N={N} K={K} {us:8.1f} us {tf:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
