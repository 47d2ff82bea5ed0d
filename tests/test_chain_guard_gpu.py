"""Robustness of the chained (counter-linked) GEMM launches: two inference plans issued on two streams
must be serialised by the host guard (engine.py::_chain_guard) and stay correct, and a chained consumer
whose producer never arrives must fail loudly (bounded spin -> watchdog record -> trap -> RuntimeError
naming the cause) instead of hanging the device."""
import os
import subprocess
import sys

import pytest
import torch

from _cases import ROOT, mel_input, noise_input
from _synth import synth_state_dict

pytestmark = pytest.mark.gpu


def _model(seed):
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    torch.manual_seed(0)
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    m.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed), strict=False)
    return m.cuda().eval()


def test_two_plans_on_two_streams_stay_correct():
    from flow2gan_b200 import _lib as L
    ma, mb = _model(11), _model(12)
    mel_a, nz_a = mel_input(16, 100, 94, seed=0).cuda(), noise_input(16, 94 * 256, seed=1).cuda()
    mel_b, nz_b = mel_input(8, 100, 61, seed=2).cuda(), noise_input(8, 61 * 256, seed=3).cuda()
    with torch.no_grad():
        for _ in range(2):                                   # eager call, then graph capture
            ya = ma.infer(mel_a, n_timesteps=2, noise=nz_a).clone()
            yb = mb.infer(mel_b, n_timesteps=1, noise=nz_b).clone()
        torch.cuda.synchronize()
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        outs = []
        for it in range(25):                                 # graph replays racing on two streams
            with torch.cuda.stream(sa):
                a = ma.infer(mel_a, n_timesteps=2, noise=nz_a)
            with torch.cuda.stream(sb):
                b = mb.infer(mel_b, n_timesteps=1, noise=nz_b)
            outs.append((a, b))
        torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, ya) and torch.equal(b, yb)
    rec = (L.C.c_int * 4)()
    assert L.lib().f2g_chain_watchdog(rec) == 0


_STARVED = r'''
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from flow2gan_b200 import _lib as L
M, C, H = 512, 384, 1152
g = torch.Generator().manual_seed(0)
a = torch.randn(M, C, generator=g).half().cuda(); W1 = torch.randn(H, C, generator=g).half().cuda()
W2 = torch.randn(C, H, generator=g).half().cuda(); b1 = torch.zeros(H).cuda(); b2 = torch.zeros(C).cuda()
sl = torch.full((H,), 0.1).cuda(); rs = torch.ones(C).cuda()
h = torch.zeros(M, H, dtype=torch.float16, device="cuda"); x = torch.zeros(M, C, device="cuda")
cnt = torch.full((2,), -1000000, dtype=torch.int32, device="cuda")       # the producer counters can never reach the target
d1 = L.gemm_desc(a.data_ptr(), W1.data_ptr(), h.data_ptr(), M, H, C, C, C, H, bias=b1.data_ptr(), slope=sl.data_ptr(),
                 act=L.ACT_PRELU, ab_f16=1, c_f16=1, done_counter=cnt.data_ptr())
d2 = L.gemm_desc(h.data_ptr(), W2.data_ptr(), x.data_ptr(), M, C, H, H, H, C, bias=b2.data_ptr(), res=x.data_ptr(), ld_res=C,
                 res_scale=rs.data_ptr(), ab_f16=1, wait_counter=cnt.data_ptr())
L.gemm_group([d1, d2])
try:
    torch.cuda.synchronize()
    print("NO ERROR")
except Exception as e:
    print("SYNC ERROR:", repr(e)[:200])
rec = (L.C.c_int * 4)()
print("WATCHDOG", L.load().f2g_chain_watchdog(rec), list(rec))
try:
    L.gemm_group([d1])
    print("NO ERROR ON NEXT CALL")
except RuntimeError as e:
    print("NEXT CALL:", str(e)[:400])
'''


def test_starved_chained_consumer_traps_with_a_report():
    r = subprocess.run([sys.executable, "-c", _STARVED.format(root=ROOT)], capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert "SYNC ERROR" in out, out[-1500:]
    assert "WATCHDOG 1" in out, out[-1500:]
    assert "timed out waiting for its producer tiles" in out, out[-1500:]
