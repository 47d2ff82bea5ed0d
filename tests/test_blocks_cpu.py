"""The generator's SIMT kernels (csrc/blocks.cu: BiasNorm, the fused ConvNeXt-block prologue, the small
dense layers, the time embedding, packing helpers) on the CPU: the same source compiled for the host
(cooperative emulation, tests/_emul.py) and called through the product's own ctypes wrappers, against
the oracle.  Mirrors tests/test_kernels_gpu.py::test_biasnorm_and_block_pre / test_time_embedding_path /
test_im2col_and_masks (reference: flow2gan/models/modules.py:217-232,286-416,456-495)."""
import math

import pytest
import torch

import _emul
from _cases import rel_rms
from oracle import flow2gan_oracle as O

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.fixture
def L(monkeypatch):
    return _emul.native_fixture(monkeypatch)


def _prologue_case(C, B, T, Tc, factor, masked, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, T, generator=g)
    bias = torch.randn(C, generator=g) * 0.1
    ls = torch.tensor(0.7)
    dw = torch.randn(C, 1, 7, generator=g) * 0.3
    dwb = torch.randn(C, generator=g) * 0.1
    cond = torch.randn(B, C, Tc, generator=g)
    zero_vec = torch.randn(C, generator=g)
    ts = torch.randn(B, C, generator=g) * 0.3
    lens = torch.tensor([T] + [max(1, T - 11)] * (B - 1))
    mask = (torch.arange(T)[None] < lens[:, None]).float() if masked else torch.ones(B, T)
    conv = torch.nn.functional.conv1d(x * mask[:, None], dw, dwb, padding=3, groups=C)
    z = O.bias_norm(conv, bias, ls)
    cup = torch.cat([cond.repeat_interleave(factor, 2),
                     zero_vec[None, :, None].expand(B, C, T - Tc * factor)], 2)
    full = ((z + cup) * (1 + ts[:, :, None])).contiguous()
    t = dict(xr=x.transpose(1, 2).reshape(B * T, C).contiguous(), dwT=dw[:, 0, :].t().contiguous(), dwb=dwb,
             bias=bias, ls=ls, mask=mask.reshape(-1).contiguous() if masked else None,
             crow=torch.cat([cond.transpose(1, 2).reshape(B * Tc, C), zero_vec[None]], 0).contiguous(), ts=ts)
    return x, t, conv, full


@pytest.mark.parametrize("C", [384, 512, 768])
def test_biasnorm_and_block_pre(L, C):
    B, T, Tc, factor = 2, 37, 9, 4
    x, t, conv, full = _prologue_case(C, B, T, Tc, factor, True, C)
    y = torch.empty_like(t["xr"])
    inv = torch.empty(B * T)
    L.biasnorm(t["xr"], B * T, C, C, t["bias"], t["ls"], y, C, inv)
    ref = O.bias_norm(x, t["bias"], t["ls"])
    assert rel_rms(y.view(B, T, C).transpose(1, 2), ref) < 2e-6
    out = torch.empty(B * T, C)
    conv_out = torch.empty(B * T, C)
    L.block_pre(t["xr"], B, T, C, C, t["dwT"], t["dwb"], t["bias"], t["ls"], t["mask"], t["crow"], C, Tc, factor,
                B * Tc, t["ts"], C, out, C, conv_out, None)
    assert rel_rms(out.view(B, T, C).transpose(1, 2), tf32_round(full)) < 3e-4          # tf32 rounding ties
    assert rel_rms(conv_out.view(B, T, C).transpose(1, 2), conv) < 2e-6
    out16 = torch.empty(B * T, C, dtype=torch.float16)
    L.block_pre(t["xr"], B, T, C, C, t["dwT"], t["dwb"], t["bias"], t["ls"], t["mask"], t["crow"], C, Tc, factor,
                B * Tc, t["ts"], C, out16, C, None, None)
    assert rel_rms(out16.float().view(B, T, C).transpose(1, 2), full) < 3e-4
    assert float((out16.float() - out).abs().max()) <= float(full.abs().max()) * 2 ** -10


def test_block_pre_group_three_branches_interior_fast_path(L):
    """The grouped launch of the three decoder branches (C = 768 / 512 / 384, frame factors 1 / 2 / 4),
    unmasked so that interior CTAs take the window fast path; the same launch clears the chaining
    counters of the GEMM group that follows."""
    B, Tc = 2, 12
    cases, descs, outs = [], [], []
    for C, factor in ((768, 1), (512, 2), (384, 4)):
        T = Tc * factor + 1                                     # one zero-padded tail frame
        x, t, conv, full = _prologue_case(C, B, T, Tc, factor, False, 100 + C)
        out = torch.full((B * T, C), float("nan"), dtype=torch.float16)
        descs.append(L.block_pre_desc(t["xr"], B, T, C, C, t["dwT"], t["dwb"], t["bias"], t["ls"], None, t["crow"], C,
                                      Tc, factor, B * Tc, t["ts"], C, out, C))
        cases.append((t, full, B, T, C))
        outs.append(out)
    counters = torch.full((40,), 7, dtype=torch.int32)
    L.block_pre_group(descs, zero=counters)
    assert int(counters.abs().sum()) == 0
    for (t, full, B_, T, C), out in zip(cases, outs):
        assert rel_rms(out.float().view(B_, T, C).transpose(1, 2), full) < 3e-4


def test_time_embedding_path(L):
    B, dim = 5, 512
    g = torch.Generator().manual_seed(0)
    t = torch.rand(B, generator=g)
    ref = O.sinusoidal_pos_emb(t, dim)
    half = dim // 2
    freqs = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    emb = torch.empty(B, dim)
    L.time_sinusoid(t, B, dim, freqs, 1000.0, emb)
    assert float((emb - ref).abs().max()) < 2e-4
    W = torch.randn(700, dim, generator=g) / math.sqrt(dim)
    bb = torch.randn(700, generator=g)
    out = torch.empty(B, 700)
    L.linear_small(emb, B, dim, dim, W, dim, bb, 700, L.ACT_SILU, out, 700)
    assert rel_rms(out, torch.nn.functional.silu(emb @ W.t() + bb)) < 2e-6


def test_im2col_masks_and_packing(L):
    B, C, T = 2, 100, 13
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, T, generator=g)
    ld = 304
    out = torch.empty(B * T, ld)
    L.im2col_cf(x, B, C, T, 3, out, ld, 0)
    xp = torch.nn.functional.pad(x, (1, 1))
    ref = torch.zeros(B, T, ld)
    for k in range(3):
        ref[:, :, k * C:(k + 1) * C] = xp[:, :, k:k + T].transpose(1, 2)
    assert torch.equal(out.view(B, T, ld), ref)
    lens = torch.tensor([1000, 777], dtype=torch.int32)
    m = torch.empty(2 * 9)
    L.frame_mask(lens, 2, 9, 128, m)
    assert torch.equal(m.view(2, 9), (torch.arange(9)[None] < (1 + lens // 128)[:, None]).float())
    with pytest.raises(RuntimeError, match="multiple of 128"):
        L.biasnorm(torch.zeros(4, 100), 4, 100, 100, torch.zeros(100), torch.zeros(()), torch.zeros(4, 100), 100)
