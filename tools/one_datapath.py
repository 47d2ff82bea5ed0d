"""Data-path / model-average kernels once at production sizes, between cudaProfilerStart/Stop, for
`ncu --set full` rows: f2g_average_update over the whole generator (79 M parameters, fp64 accumulator),
f2g_gain_resample (16 x 4 s, 44.1 -> 24 kHz), f2g_pcm16_encode, f2g_pcm_decode."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from flow2gan_b200.averaging import average_state_dict
from flow2gan_b200 import datapath as DP
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
avg = {k: v.detach().double().clone() for k, v in m.state_dict().items()}
cur = m.state_dict()
x = (torch.randn(16, 4 * 44100, device=dev) * 0.1).clamp(-1, 1)
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    average_state_dict(avg, cur, 0.99, 0.01)
    y = torch.stack([DP.gain_resample(x[i].contiguous(), 44100, 24000) for i in range(2)])
    pcm = DP.encode_pcm16(y)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ok", y.shape, pcm.shape)
