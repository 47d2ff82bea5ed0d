"""One fp16 CTA-pair GEMM launch at a steady-state size, for an ncu --set full --import-source capture:
   ncu ... -k regex:gemm_pair -s 3 -c 1 python tools/one_gemm.py [h|res] [K] [waves]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flow2gan_b200 import _lib as L
L.lib()
epi = sys.argv[1] if len(sys.argv) > 1 else "h"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 384
waves = int(sys.argv[3]) if len(sys.argv) > 3 else 4
M, N = 256 * 74 * waves // 4, 1024
a = torch.randn(M, K, device="cuda").half(); b = (torch.randn(N, K, device="cuda") * 0.02).half()
bias = torch.randn(N, device="cuda"); sl = torch.rand(N, device="cuda")
if epi == "h":
    c = torch.empty(M, N, device="cuda", dtype=torch.float16)
    kw = dict(bias=bias.data_ptr(), slope=sl.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1)
else:
    c = torch.zeros(M, N, device="cuda")
    kw = dict(bias=bias.data_ptr(), res=c.data_ptr(), ld_res=N, res_scale=sl.data_ptr(), ab_f16=1)
d = L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, K, N, **kw)
for _ in range(5):
    L.gemm_group([d])
torch.cuda.synchronize()
