"""One pwconv1 group + one pwconv2 group launch (bench shape) for ncu source-level captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import pair_bench as pb
g1, g2, keep = pb.block_groups()
from flow2gan_b200 import _lib as L
for _ in range(2):
    L.gemm_group(g1); L.gemm_group(g2)
torch.cuda.synchronize()
