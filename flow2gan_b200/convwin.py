"""Windowed ("implicit im2col") Conv2d for the discriminators: stride (1, sw) convolutions on
channel-last tensors whose GEMM A operand is read by TMA straight from a zero-padded copy of the
input through an OVERLAPPING-row tensor map -- the (kh*kw/sw)-times larger im2col matrix of the
gather path (discriminators.py::_Conv2dCLFn) is never written, and the input gradient is a
phase-decomposed transposed convolution (sw gather-GEMMs, no col2im scatter, no atomics).

Reference call sites: torch.nn.Conv2d of DiscriminatorP ((5,1) stride (3,1), run here along W
on a (B*period, 1, T/period, C) tensor) and DiscriminatorR ((3,9) stride (1,2) / (3,3)),
flow2gan/models/discriminators.py:65-76,95-104,171-184,203-217.

Geometry (all sizes in elements; C = input channels, multiple of 32):

  xp  : (Nb, Hl, Wp, C) zero-padded input, Hl = H + 2*ph lines, Wp = sw*R columns per line,
        data at [ph, ph+H) x [pw, pw+W).  R = output columns ALLOCATED per line (>= Wo).
  y   : (Nb*Hl*R, Cop) = one GEMM row per (n, line, r); valid outputs are line < Ho, r < Wo, the
        rest ("garbage rows") is computed from padding / neighbouring lines and never read.
  fwd : A row m starts at xp + m*sw*C, segment s (kernel row) is kw*C long and lives R rows
        further:  y[m, :] = sum_s  xp_row(m + s*R)[0 : kw*C] . Wf[:, s*kw*C : (s+1)*kw*C]^T
  wgrad: dWt[(s, t, ci), co] = sum_m xp_row(m + s*R)[t*C + ci] * dz[m, co]   (A = xp, MN-major)
  dgrad: dxp[(line, r), phase*C + ci] = sum_{s, j, co} dz[(line - s, r - j), co] * W[co, ci, s, phase + sw*j]
        one GEMM per phase in [0, sw): A row m' = window of nt_phase consecutive dz rows ending
        at m' - s*R; garbage rows of dz are zero, R - Wo >= nt - 1 keeps the windows that
        straddle a line start on them (and a zero guard precedes row 0).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
from torch import Tensor

from . import _lib as L


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b


@dataclass(frozen=True)
class WinGeom:
    Nb: int
    H: int
    W: int
    C: int
    Co: int
    kh: int
    kw: int
    sw: int
    ph: int
    pw: int
    Ho: int
    Wo: int
    Hl: int
    R: int
    Wp: int
    Cop: int
    nt: int          # max taps per phase = ceil(kw / sw)
    M: int           # GEMM rows = Nb * Hl * R
    K: int           # kh * kw * C
    seg: int         # kw * C
    guard: int       # zero rows in front of dz (nt - 1)

    def taps(self, phase: int) -> int:
        return _ceil(self.kw - phase, self.sw) if phase < self.kw else 0


def win_geom(Nb: int, H: int, W: int, C: int, Co: int, kh: int, kw: int, sw: int, ph: int, pw: int) -> WinGeom:
    assert C % 32 == 0 and sw >= 1 and kw >= sw, (C, sw, kw)
    Ho = H + 2 * ph - kh + 1
    Wo = (W + 2 * pw - kw) // sw + 1
    assert Ho >= 1 and Wo >= 1
    nt = _ceil(kw, sw)
    R = max(_ceil(W + 2 * pw, sw), Wo + nt - 1)
    Hl = H + 2 * ph
    return WinGeom(Nb, H, W, C, Co, kh, kw, sw, ph, pw, Ho, Wo, Hl, R, sw * R, _ceil(Co, 32) * 32, nt,
                   Nb * Hl * R, kh * kw * C, kw * C, nt - 1)


def supports(C: int, kh: int, kw: int, sh: int, sw: int) -> bool:
    return C % 32 == 0 and sh == 1 and kw >= sw


def pack_fwd_weight(weight: Tensor, g: WinGeom) -> Tensor:
    """(Co, Ci, kh, kw) -> (Cop, K) with k = (s, t, ci); rows >= Co are zero."""
    wf = weight.new_zeros(g.Cop, g.K)
    wf[:g.Co] = weight.permute(0, 2, 3, 1).reshape(g.Co, g.K)
    return wf


def pack_dgrad_weight(weight: Tensor, g: WinGeom, phase: int) -> Tensor:
    """(Co, Ci, kh, kw) -> (Ci, kh * nt_phase * Cop): column (s, jj, co) holds
    W[co, ci, s, phase + sw*(nt_phase-1-jj)] (the window runs over ascending dz rows)."""
    ntp = g.taps(phase)
    sel = weight[:, :, :, phase::g.sw].flip(-1)                   # (Co, Ci, kh, ntp), jj ascending
    if g.Cop == g.Co:
        wd = sel.permute(1, 2, 3, 0).contiguous()
    else:
        wd = weight.new_zeros(g.C, g.kh, ntp, g.Cop)
        wd[..., :g.Co] = sel.permute(1, 2, 3, 0)
    return wd.view(g.C, g.kh * ntp * g.Cop)


def _round_inplace(w: Tensor) -> None:
    rows, cols = w.shape
    L.pack2d(w.data_ptr(), cols, 1, rows, cols, w.data_ptr(), cols, cols, 1)


class _ConvWinFn(torch.autograd.Function):
    """y = leaky_relu(conv2d(x, stride (1, sw)) + b) on channel-last (Nb, H, W, C) ->
    (Nb, Ho, Wo, Co); the result is a strided view of the (Nb, Hl, R, Cop) GEMM output."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, sw: int, ph: int, pw: int, leaky: Optional[float],
                packs: Optional[dict] = None):
        """`packs` (optional): a per-module cache dict owned by the caller (discriminators.conv2d_cl): the
        GEMM-ready weight layouts ("wf" forward, "wd" transposed-convolution) are then (re)built only when
        the weight tensor's version changed, in place, instead of on every call."""
        Nb, H, W, Cc = x.shape
        Co, Ci, kh, kw = weight.shape
        assert Ci == Cc and x.stride(3) == 1, (x.shape, x.stride())
        g = win_geom(Nb, H, W, Cc, Co, kh, kw, sw, ph, pw)
        dev = x.device
        xp = torch.empty(g.M * sw * Cc + g.seg, device=dev, dtype=torch.float32)
        L.pad2d(x.data_ptr(), Nb, H, W, Cc, x.stride(0), x.stride(1), x.stride(2), g.Hl, g.Wp, ph, pw, g.seg, xp)
        wf = packs.get("wf") if packs is not None else None
        if wf is None or packs.get("wf_ver") != weight._version:
            if wf is None:
                wf = torch.empty(g.Cop, g.K, device=dev, dtype=torch.float32)
            L.conv_w_pack(weight.detach().contiguous(), Co, Ci, kh * kw, g.Cop, g.K, wf, 0)
            if packs is not None:
                packs["wf"], packs["wf_ver"] = wf, weight._version
        y = torch.empty(g.M, g.Cop, device=dev, dtype=torch.float32)
        if g.Cop != Co:
            y.zero_()
        L.gemm_group([L.gemm_desc(xp.data_ptr(), wf.data_ptr(), y.data_ptr(), g.M, Co, g.K, sw * Cc, g.K, g.Cop,
                                  bias=bias.data_ptr(), act=L.ACT_LEAKY if leaky is not None else L.ACT_NONE,
                                  leaky=leaky or 0.0, a_seg_len=g.seg, a_seg_shift=g.R, a_rows=g.M)])
        ctx.g, ctx.leaky, ctx.packs = g, leaky, packs
        need_w = ctx.needs_input_grad[1]
        ctx.saved = (xp if need_w else None, y, weight.detach())
        return y.view(Nb, g.Hl, g.R, g.Cop)[:, :g.Ho, :g.Wo, :Co]

    @staticmethod
    def backward(ctx, dy: Tensor):
        g: WinGeom = ctx.g
        leaky = ctx.leaky
        xp, y, weight = ctx.saved
        dev = dy.device
        Cop, C, sw = g.Cop, g.C, g.sw
        need_w = ctx.needs_input_grad[1]
        # one zero-filled buffer for everything that is accumulated into (bias gradient, split-K
        # weight gradient): one fill launch instead of two
        zb = torch.zeros(Cop + (g.K * Cop if need_w else 0), device=dev, dtype=torch.float32)
        g_bias = zb[:Cop]
        act = L.ACT_LEAKY if leaky is not None else L.ACT_NONE
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        fused = g.Co % 4 == 0 and all(st % 4 == 0 for st in dy.stride()[:3]) and dy.data_ptr() % 16 == 0
        if fused:
            # gather dy, apply act', zero the padding rows / guard, bias sums: ONE pass (f2g_act_bwd_win)
            dz_full = torch.empty((g.guard + g.M) * Cop, device=dev, dtype=torch.float32)
            dz = dz_full[g.guard * Cop:].view(g.M, Cop)
            L.act_bwd_win(dy, g.Nb, g.Hl, g.R, g.Ho, g.Wo, y if leaky is not None else None, Cop, leaky or 0.0,
                          act, g.Co, Cop, dz.data_ptr(), Cop, g.guard, g_bias)
        else:
            dz_full = torch.zeros((g.guard + g.M) * Cop, device=dev, dtype=torch.float32)
            dz = dz_full[g.guard * Cop:].view(g.M, Cop)
            dz.view(g.Nb, g.Hl, g.R, Cop)[:, :g.Ho, :g.Wo, :g.Co].copy_(dy)
            L.act_bwd(dz, Cop, y if leaky is not None else None, Cop, None, leaky or 0.0, act, g.M, g.Co, dz, Cop,
                      g_bias, None, round_tf32=1)
        gW = gb = gx = None
        if ctx.needs_input_grad[1]:
            dwt = zb[Cop:].view(g.K, Cop)
            L.gemm_group([L.gemm_desc(xp.data_ptr(), dz.data_ptr(), dwt.data_ptr(), g.K, g.Co, g.M, sw * C, Cop, Cop,
                                      a_mn=1, b_mn=1, split_k=L.pick_split_k(g.K, g.Co, g.M),
                                      a_seg_len=g.seg, a_seg_shift=g.R, a_rows=g.M)])
            gW = dwt.view(g.kh, g.kw, C, Cop)[..., :g.Co].permute(3, 2, 0, 1).contiguous()
        if ctx.needs_input_grad[2]:
            gb = g_bias[:g.Co]
        if ctx.needs_input_grad[0]:
            dxp = torch.empty(g.M, sw * C, device=dev, dtype=torch.float32)
            descs: List[L.F2GGemm] = []
            keep = []
            # all phases' transposed-conv weights in one launch, phase blocks back to back (cached per
            # weight version when the caller supplied a cache)
            packs = ctx.packs
            wd_all = packs.get("wd") if packs is not None else None
            if wd_all is None or packs.get("wd_ver") != weight._version:
                if wd_all is None:
                    wd_all = torch.empty(C * g.kh * g.kw * Cop, device=dev, dtype=torch.float32)
                L.conv_w_pack_dgrad(weight.contiguous(), g.Co, C, g.kh, g.kw, sw, Cop, wd_all)
                if packs is not None:
                    packs["wd"], packs["wd_ver"] = wd_all, weight._version
            keep.append(wd_all)
            off = 0
            for phase in range(sw):
                ntp = g.taps(phase)
                kd = g.kh * ntp * Cop
                a_ptr = dz_full.data_ptr() + 4 * (g.guard - (ntp - 1)) * Cop
                descs.append(L.gemm_desc(a_ptr, wd_all.data_ptr() + 4 * off, dxp.data_ptr() + 4 * phase * C, g.M, C,
                                         kd, Cop, kd, sw * C,
                                         a_seg_len=ntp * Cop, a_seg_shift=-g.R, a_rows=g.M))
                off += C * kd
            L.gemm_group(descs)
            gx = dxp.view(g.Nb, g.Hl, g.Wp, C)[:, g.ph:g.ph + g.H, g.pw:g.pw + g.W, :]
        return gx, gW, gb, None, None, None, None, None


def conv2d_win(x: Tensor, weight: Tensor, bias: Tensor, sw: int, ph: int, pw: int, leaky: Optional[float],
               packs: Optional[dict] = None) -> Tensor:
    return _ConvWinFn.apply(x, weight, bias, sw, ph, pw, leaky, packs)
