"""Bring-up helper: sweeps MN-major UMMA descriptor geometries (F2G_MN_* env overrides) on
one small GEMM per operand-major combination and prints the rel-RMS error of each variant."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from flow2gan_b200 import _lib as L
from test_kernels_gpu import tf32_round
from _cases import rel_rms
L.lib()
for a_mn, b_mn in ((0,1),(1,0),(1,1)):
    M,N,K = 256,128,64
    g = torch.Generator().manual_seed(0)
    A = tf32_round(torch.randn(M,K,generator=g)); B = tf32_round(torch.randn(N,K,generator=g))
    Ad = (A.t().contiguous() if a_mn else A).cuda(); Bd = (B.t().contiguous() if b_mn else B).cuda()
    C = torch.zeros(M,N,device="cuda")
    try:
        L.gemm_group([L.gemm_desc(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), M,N,K, Ad.shape[1], Bd.shape[1], N, bn=128, a_mn=a_mn, b_mn=b_mn)])
        torch.cuda.synchronize()
        print("  a_mn=%%d b_mn=%%d err=%%.3e" %% (a_mn,b_mn, rel_rms(C.cpu(), A.double()@B.double().t())))
    except Exception as e:
        print("  a_mn=%%d b_mn=%%d EXC %%s" %% (a_mn,b_mn, str(e)[:80])); break
''' % (ROOT, ROOT)

VARIANTS = [
    dict(),                                                          # BASE32B, sbo 512, lbo 4096, TMA ATOM_32B
    dict(F2G_MN_LBO="512", F2G_MN_SBO="4096"),
    dict(F2G_MN_TMA_SWIZZLE="3"),
    dict(F2G_MN_LAYOUT="2", F2G_MN_SBO="1024", F2G_MN_TMA_SWIZZLE="3"),
    dict(F2G_MN_LAYOUT="2", F2G_MN_SBO="4096", F2G_MN_LBO="1024", F2G_MN_TMA_SWIZZLE="3"),
    dict(F2G_MN_SBO="1024"),
    dict(F2G_MN_TMA_SWIZZLE="5"),
]
for v in VARIANTS:
    env = dict(os.environ, **v)
    print("variant", v, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr[-300:] if r.returncode else "", flush=True)
