"""fp16-operand CTA-pair GEMM: steady-state time per wave of identical 256x256xK tiles and launch
intercept, for the two block epilogues (h: bias+PReLU -> fp16; res: bias+scale*residual -> fp32 in place),
optionally with the epilogue's stores (F2G_PAIR_DBG=1) or its TMEM drain too (=3) switched off -- separates
the main loop's operand stream from the epilogue's cost.   python tools/pair_f16_scan.py   (B200)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()
dev = "cuda"
SM_HZ = 1.965e9


def timeit(descs, reps=20):
    for _ in range(3): L.gemm_group(descs)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): L.gemm_group(descs)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def prob(M, N, K, epi):
    a = (torch.randn(M, K, device=dev)).half(); b = (torch.randn(N, K, device=dev) * 0.02).half()
    bias = torch.randn(N, device=dev); sl = torch.rand(N, device=dev)
    if epi == "h":
        c = torch.empty(M, N, device=dev, dtype=torch.float16)
        kw = dict(bias=bias.data_ptr(), slope=sl.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1)
    else:
        c = torch.zeros(M, N, device=dev)
        kw = dict(bias=bias.data_ptr(), res=c.data_ptr(), ld_res=N, res_scale=sl.data_ptr(), ab_f16=1)
    return L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, K, N, **kw), (a, b, c, bias, sl)


print("dbg =", os.environ.get("F2G_PAIR_DBG", "0"), flush=True)
EPIS = tuple(os.environ.get("SCAN_EPIS", "h,res").split(","))
KS = tuple(int(k) for k in os.environ.get("SCAN_KS", "384,768,1152,2304").split(","))
for epi in EPIS:
    for K in KS:
        row = []
        for waves in (1, 2, 4, 8):
            d, keep = prob(256 * 74 * waves // 4, 1024, K, epi)      # 74*waves tiles of 256x256
            row.append(timeit([d]))
        slope = (row[3] - row[2]) / 4
        stage_us = slope / (K / 64)
        print(f"{epi:3s} K={K:4d} waves 1,2,4,8: " + " ".join(f"{u:7.2f}" for u in row) +
              f"   slope {slope:6.2f} us/wave = {stage_us*1e3:5.0f} ns/stage = {32768 / (stage_us * 1e-6 * SM_HZ):5.1f} B/clk/SM operand ingest, "
              f"{2*256*256*K*74/slope/1e6:6.0f} TF/s;  intercept {row[0] - slope:6.2f} us", flush=True)
