"""Timing probe: pwconv1 group + pwconv2 group as two launches vs ONE launch of all six problems
(no dependency enforcement -- timing only), both from a CUDA graph.  python tools/merge_probe.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from flow2gan_b200 import _lib as L
L.lib()
import pair_bench as PB


def graph_time(fn, reps=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


g1, g2, keep = PB.block_groups_f16()
print("generic=%s" % os.environ.get("F2G_PAIR_FORCE_GENERIC", "0"))
print("g1 alone      %.2f us" % graph_time(lambda: L.gemm_group(g1)))
print("g2 alone      %.2f us" % graph_time(lambda: L.gemm_group(g2)))
print("g1 ; g2       %.2f us" % graph_time(lambda: (L.gemm_group(g1), L.gemm_group(g2))))
print("g1 + g2 fused %.2f us" % graph_time(lambda: L.gemm_group(g1 + g2)))
