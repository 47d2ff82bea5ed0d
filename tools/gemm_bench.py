"""Micro-benchmark of the grouped TF32 GEMM (bring-up / tuning).  Usage:
   python tools/gemm_bench.py [M N K bn]...   (env F2G_GEMM_DBG for pipeline experiments)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()

def run(M, N, K, bn, reps=20, a_mn=0, b_mn=0):
    A = torch.randn(K, M, device="cuda") if a_mn else torch.randn(M, K, device="cuda")
    B = torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")
    C = torch.empty(M, N, device="cuda")
    d = L.gemm_desc(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, A.shape[1], B.shape[1], N, bn=bn, a_mn=a_mn, b_mn=b_mn)
    for _ in range(3): L.gemm_group([d])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): L.gemm_group([d])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    tiles = -(-M // 128) * -(-N // bn)
    print(f"dbg={os.environ.get('F2G_GEMM_DBG','0')} M={M} N={N} K={K} bn={bn} mn=({a_mn},{b_mn}) tiles={tiles} {us:8.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    cases = [(1520, 2304, 768, 128), (1520, 2304, 768, 256), (1520, 768, 2304, 64), (1520, 768, 2304, 128),
             (18944, 2304, 768, 128), (18944, 2304, 768, 256), (128, 128, 4096, 128), (128, 256, 4096, 256),
             (18944, 128, 768, 128)]
    for c in cases: run(*c)
    run(2304, 768, 1520, 128, a_mn=1, b_mn=1)
    run(1520, 768, 2304, 128, b_mn=1)
