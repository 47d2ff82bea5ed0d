"""Differentiable generator path: forward with saved activations + explicit backward, both as
launch sequences of the CUDA library, exposed to autograd as two torch.autograd.Functions
(CondEncoder and one model evaluation `process_model`).  Used by MelAudioGenerator.forward
(stage-1 flow-matching loss, flow2gan/models/generator.py:172-234,294-325) and by
MelAudioGenerator.infer under grad (GAN G-phase, flow2gan/models/gan.py:133-143).

Gradient formulas follow the reference's autograd of modules.py (ConvNeXtBlock.forward :473-495,
BiasNormFunction :286-339, upsample_cond :668-680, STFT/ISTFT :69-116); contractions are
tcgen05 TF32 GEMMs (dgrad = K-major x MN-major, wgrad = MN-major x MN-major), everything else
SIMT kernels from csrc/train.cu / csrc/spectral.cu.
"""
from __future__ import annotations

import random
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib as L
from .engine import BN_G1, BN_G2, PackedGenerator


# Zero-initialised accumulators of one backward call (bias / slope / norm-parameter gradients, split-K
# weight gradients ...: ~400 small tensors per generator backward) are carved out of ONE zero-filled
# buffer instead of one fill launch each.  The buffer is allocated per backward call and never reused:
# gradients that autograd keeps (views of it) stay valid for as long as they are referenced.
_ARENA = None            # [flat fp32 tensor, next free element] while a backward call is running
_ARENA_FLOATS = 3 << 20      # 12 MB: one fill of ~2 us instead of ~400 x 1.3 us
_ARENA_MAX_ITEM = 1 << 18


class _zero_arena:
    def __init__(self, dev):
        self.dev = dev

    def __enter__(self):
        global _ARENA
        self.prev = _ARENA
        _ARENA = [torch.zeros(_ARENA_FLOATS, device=self.dev, dtype=torch.float32), 0]

    def __exit__(self, *exc):
        global _ARENA
        _ARENA = self.prev
        return False


def _z(*shape, dev):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    n = 1
    for d in shape:
        n *= int(d)
    a = _ARENA
    if a is not None and 0 < n <= _ARENA_MAX_ITEM and a[0].device == torch.device(dev):
        start = (a[1] + 63) & ~63                       # 256-byte aligned slices
        if start + n <= a[0].numel():
            a[1] = start + n
            return a[0][start:start + n].view(tuple(int(d) for d in shape))
    return torch.zeros(*shape, device=dev, dtype=torch.float32)


def _e(*shape, dev):
    return torch.empty(*shape, device=dev, dtype=torch.float32)


def _nt(a, lda, b, ldb, c, ldc, M, N, K, bn=128, **kw):
    """C = A[M,K] . B[N,K]^T, both K-major (forward)."""
    return L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, lda, ldb, ldc, bn=bn, **kw)


def _nn(a, lda, b, ldb, c, ldc, M, N, K, bn=128, **kw):
    """dgrad: C[M,N] = A[M,K] . Bs[K,N]   (A K-major, Bs stored (K, ldb) = MN-major B)."""
    return L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, lda, ldb, ldc, bn=bn, b_mn=1, **kw)


def _tn(a, lda, b, ldb, c, ldc, M, N, K, bn=128, **kw):
    """wgrad: C[M,N] = As[K,M]^T . Bs[K,N]   (both stored row-major over K = MN-major)."""
    return L.gemm_desc(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, lda, ldb, ldc, bn=bn, a_mn=1,
                       b_mn=1, **kw)


def _limit_flip(grad: Tensor, x: Tensor, lo: float, hi: float) -> Tensor:
    """LimitParamValue.backward (modules.py:245-256) applied to one call's gradient."""
    g = grad * torch.where((grad > 0) & (x < lo), -1.0, 1.0)
    return g * torch.where((g < 0) & (x > hi), -1.0, 1.0)


class DrawBuffer:
    """Source of the LimitParamValue coin flips (modules.py:267) when a whole training iteration
    is captured in a CUDA graph: the forward pass takes device scalars out of `dev` instead of
    calling random.random(), and the host refills them -- same generator, same order, same
    count as the eager path -- before every replay."""

    def __init__(self, device, n: int = 512):
        self.dev = torch.zeros(n, device=device)
        self.host = torch.zeros(n).pin_memory()
        self.i = 0
        self.count = 0

    def take(self) -> Tensor:
        t = self.dev[self.i]
        self.i += 1
        return t

    def refill(self, prob: float = 0.6) -> None:
        for k in range(self.count):
            self.host[k] = 1.0 if random.random() < prob else 0.0
        if self.count:
            self.dev[: self.count].copy_(self.host[: self.count], non_blocking=True)


_draws: Optional[DrawBuffer] = None      # set by GANTrainer while it captures a step graph


def _limit_draw(training: bool, prob: float = 0.6):
    if not training:
        return False
    if _draws is not None:
        return _draws.take()
    return random.random() < prob          # modules.py:267


def _limit_apply(flag, grad: Tensor, x: Tensor, lo: float, hi: float) -> Tensor:
    """flag: Python bool (eager) or a device 0/1 scalar (graph mode, no host decision)."""
    if isinstance(flag, Tensor):
        return torch.where(flag > 0, _limit_flip(grad, x, lo, hi), grad)
    return _limit_flip(grad, x, lo, hi) if flag else grad


# =========================================================================================
# ConvNeXt block stack (shared by the cond encoder and the three decoders)
# =========================================================================================
class _BlockSave:
    __slots__ = ("x", "y", "inv", "a1", "hpre", "h", "lim_norm", "lim_rs")


def _blocks_forward(blocks_w, x: Tensor, B: int, T: int, C: int, mask, cond, ld_cond, cond_T, factor,
                    zero_row, ts, training: bool, round_last: bool):
    """Runs the blocks on x (B*T, C); returns (x_out, saves).  cond/ts: (tensor, col stride) or None."""
    dev = x.device
    R = B * T
    saves = []
    nl = len(blocks_w)
    for i, bw in enumerate(blocks_w):
        b = bw.blk
        s = _BlockSave()
        s.x = x
        s.y = _e(R, C, dev=dev)
        s.inv = _e(R, dev=dev)
        s.a1 = _e(R, C, dev=dev)
        s.hpre = _e(R, bw.H, dev=dev)
        s.h = _e(R, bw.H, dev=dev)
        s.lim_norm = _limit_draw(training)
        s.lim_rs = _limit_draw(training)
        L.block_pre(x, B, T, C, C, bw.dwT, b.dwconv.bias, b.norm.bias, b.norm.log_scale, mask,
                    None if cond is None else cond[:, i * C:], ld_cond, cond_T, factor, zero_row,
                    None if ts is None else ts[:, i * C:], 0 if ts is None else ts.stride(0),
                    s.a1, C, s.y, s.inv)
        L.gemm_group([_nt(s.a1, C, bw.W1, C, s.h, bw.H, R, bw.H, C, bn=BN_G1, bias=b.pwconv1.bias.data_ptr(),
                          slope=b.act.weight.data_ptr(), act=L.ACT_PRELU, round_tf32=1,
                          c_pre=s.hpre.data_ptr(), ld_pre=bw.H)])
        xn = _e(R, C, dev=dev)
        L.gemm_group([_nt(s.h, bw.H, bw.W2, bw.H, xn, C, R, C, bw.H, bn=BN_G2, bias=b.pwconv2.bias.data_ptr(),
                          res=x.data_ptr(), ld_res=C, res_scale=b.residual_scale.scale.data_ptr(),
                          round_tf32=int(round_last and i == nl - 1))])
        saves.append(s)
        x = xn
    return x, saves


def _blocks_backward(blocks_w, saves, dxo: Tensor, B: int, T: int, C: int, mask, cond, ld_cond, cond_T,
                     factor, zero_row, ts, g_ts, dcond, grads: Dict[str, Tensor], prefix: str):
    """dxo: grad w.r.t. the stack output.  Fills grads[prefix + 'blocks.i....'], accumulates the
    time-scale grads into g_ts (B, nl*C) and writes cond-row grads into dcond (Rc, nl*C)."""
    dev = dxo.device
    R = B * T
    for i in reversed(range(len(blocks_w))):
        bw, s = blocks_w[i], saves[i]
        b = bw.blk
        H = bw.H
        pre = f"{prefix}blocks.{i}."
        dh = _e(R, H, dev=dev)
        L.gemm_group([_nn(dxo, C, bw.W2, H, dh, H, R, H, C, bn=BN_G1)])
        g_b1, g_sl = _z(H, dev=dev), _z(H, dev=dev)
        L.act_bwd(dh, H, s.hpre, H, b.act.weight, 0.0, L.ACT_PRELU, R, H, dh, H, g_b1, g_sl, round_tf32=1)
        sk = L.pick_split_k(C, H, R)                # few output tiles, long K: split over the CTA pairs
        gW2 = _z(C, H, dev=dev) if sk > 1 else _e(C, H, dev=dev)
        gW1 = _z(H, C, dev=dev) if sk > 1 else _e(H, C, dev=dev)
        L.gemm_group([_tn(dxo, C, s.h, H, gW2, H, C, H, R, split_k=sk)])
        L.gemm_group([_tn(dh, H, s.a1, C, gW1, C, H, C, R, split_k=sk)])
        da1 = _e(R, C, dev=dev)
        L.gemm_group([_nn(dh, H, bw.W1, C, da1, C, R, C, H, bn=BN_G2)])
        dy, coef, gs = _e(R, C, dev=dev), _e(R, dev=dev), _e(R, dev=dev)
        du = _e(R, C, dev=dev) if cond is not None else None
        ts_i = None if ts is None else ts[:, i * C:]
        L.block_bwd_a(da1, C, s.y, s.inv, b.norm.bias, b.norm.log_scale, ts_i,
                      0 if ts is None else ts.stride(0), B, T, C, dy, du, coef, gs)
        g_dww, g_dwb = _z(7, C, dev=dev), _z(C, dev=dev)
        g_beta, g_ls = _z(C, dev=dev), _z((), dev=dev)
        g_rs, g_b2 = _z(C, dev=dev), _z(C, dev=dev)
        kw = dict(dy=dy, x=s.x, ld_x=C, row_mask=mask, y=s.y, coef=coef, gs=gs, bn_bias=b.norm.bias,
                  da1=da1, ld_da=C, inv=s.inv, dxo=dxo, ld_dxo=C, g_dww=g_dww, g_dwb=g_dwb,
                  g_beta=g_beta, g_ls=g_ls, g_rs=g_rs, g_b2=g_b2, B=B, T=T, C=C)
        if cond is not None:
            kw.update(cond=cond[:, i * C:], ld_cond=ld_cond, cond_T=cond_T, factor=factor, zero_row=zero_row)
        if ts is not None:
            kw.update(g_ts=g_ts[:, i * C:], ld_gts=g_ts.stride(0))
        L.block_bwd_c(**kw)
        if cond is not None:
            L.cond_reduce(du, B, T, C, cond_T, factor, zero_row, dcond[:, i * C:], dcond.stride(0))
        dx = _e(R, C, dev=dev)
        L.block_bwd_b(dy, bw.dwT, mask, dxo, C, b.residual_scale.scale, B, T, C, dx, C)
        gdw = _e(C, 1, 7, dev=dev)
        L.pack2d(g_dww.data_ptr(), 1, C, C, 7, gdw.data_ptr(), 7, 7, 0)
        g_ls = _limit_apply(s.lim_norm, g_ls, b.norm.log_scale.detach(), -1.5, 1.5)
        g_rs = _limit_apply(s.lim_rs, g_rs.view(C, 1), b.residual_scale.scale.detach(), 0.5, 1.0)
        grads[pre + "dwconv.weight"] = gdw
        grads[pre + "dwconv.bias"] = g_dwb
        grads[pre + "norm.bias"] = g_beta
        grads[pre + "norm.log_scale"] = g_ls
        grads[pre + "pwconv1.weight"] = gW1.view(H, C, 1)
        grads[pre + "pwconv1.bias"] = g_b1
        grads[pre + "act.weight"] = g_sl
        grads[pre + "pwconv2.weight"] = gW2.view(C, H, 1)
        grads[pre + "pwconv2.bias"] = g_b2
        grads[pre + "residual_scale.scale"] = g_rs
        dxo = dx
    return dxo


def _norm_backward(dz: Tensor, x0: Tensor, inv: Tensor, norm, R: int, C: int, B: int, T: int, lim: bool,
                   grads: Dict[str, Tensor], prefix: str) -> Tensor:
    """Stand-alone BiasNorm backward (in_norm): returns grad w.r.t. its input."""
    dev = dz.device
    dx, coef, gs = _e(R, C, dev=dev), _e(R, dev=dev), _e(R, dev=dev)
    L.block_bwd_a(dz, C, x0, inv, norm.bias, norm.log_scale, None, 0, B, T, C, dx, None, coef, gs)
    g_beta, g_ls = _z(C, dev=dev), _z((), dev=dev)
    L.block_bwd_c(y=x0, coef=coef, gs=gs, bn_bias=norm.bias, g_beta=g_beta, g_ls=g_ls, B=B, T=T, C=C)
    g_ls = _limit_apply(lim, g_ls, norm.log_scale.detach(), -1.5, 1.5)
    grads[prefix + "bias"] = g_beta
    grads[prefix + "log_scale"] = g_ls
    return dx


# =========================================================================================
# CondEncoder  (modules.py:523-542)
# =========================================================================================
class _CondEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, mel: Tensor, *params):
        pk: PackedGenerator = model.packed()
        ce = model.cond_encoder
        dev = mel.device
        B, nm, Fm = mel.shape
        M, Cc = B * Fm, ce.channels
        training = model.training
        mel_cl = _e(M, pk.ld_mel, dev=dev)
        L.im2col_cf(mel.contiguous(), B, nm, Fm, 3, mel_cl, pk.ld_mel, 1)
        x0 = _e(M, Cc, dev=dev)
        L.gemm_group([_nt(mel_cl, pk.ld_mel, pk.ce_Win, pk.ld_mel, x0, Cc, M, Cc, 3 * nm,
                          bias=ce.in_proj.bias.data_ptr())])
        x = _e(M, Cc, dev=dev)
        inv0 = _e(M, dev=dev)
        lim0 = _limit_draw(training)
        L.biasnorm(x0, M, Cc, Cc, ce.in_norm.bias, ce.in_norm.log_scale, x, Cc, inv0)
        xo, saves = _blocks_forward(pk.ce_blocks, x, B, Fm, Cc, None, None, 0, 0, 1, 0, None, training,
                                    round_last=True)
        c0 = _z(M + 1, Cc, dev=dev)                    # + the all-zero conditioning row
        c0[:M].copy_(xo)
        ctx.model, ctx.saved = model, (mel_cl, x0, inv0, lim0, saves, B, Fm)
        return c0

    @staticmethod
    def backward(ctx, dc0: Tensor):
        with _zero_arena(dc0.device):
            return _CondEncoderFn._backward(ctx, dc0)

    @staticmethod
    def _backward(ctx, dc0: Tensor):
        model = ctx.model
        pk: PackedGenerator = model._packed
        ce = model.cond_encoder
        mel_cl, x0, inv0, lim0, saves, B, Fm = ctx.saved
        M, Cc = B * Fm, ce.channels
        grads: Dict[str, Tensor] = {}
        dxo = dc0[:M].contiguous()
        dx = _blocks_backward(pk.ce_blocks, saves, dxo, B, Fm, Cc, None, None, 0, 0, 1, 0, None, None,
                              None, grads, "cond_encoder.")
        dx0 = _norm_backward(dx, x0, inv0, ce.in_norm, M, Cc, B, Fm, lim0, grads, "cond_encoder.in_norm.")
        dev = dx0.device
        nm = ce.cond_dim
        gWp = _e(Cc, pk.ld_mel, dev=dev)
        L.gemm_group([_tn(dx0, Cc, mel_cl, pk.ld_mel, gWp, pk.ld_mel, Cc, 3 * nm, M)])
        # packed column k*nm + ci -> parameter layout (co, ci, k): one tiny (150 k element) permute
        gW = gWp[:, : 3 * nm].reshape(Cc, 3, nm).permute(0, 2, 1).contiguous()
        gb = _z(Cc, dev=dev)
        L.colsum(dx0, Cc, M, Cc, gb)
        grads["cond_encoder.in_proj.weight"] = gW
        grads["cond_encoder.in_proj.bias"] = gb
        names = ctx.model._ce_param_names
        return (None, None) + tuple(grads.get(n) for n in names)


# =========================================================================================
# one model evaluation: process_model (generator.py:129-170) on encoded conditioning rows
# =========================================================================================
class _ProcessModelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, c0: Tensor, x_audio: Tensor, t_vec: Tensor, lens, weight, *params):
        pk: PackedGenerator = model.packed()
        dev = x_audio.device
        B, T = x_audio.shape
        Rc = c0.shape[0]
        Fm = (Rc - 1) // B
        training = model.training
        x_audio = x_audio.contiguous()
        dec0 = model.estimators[0].decoder
        emb = _e(B, dec0.time_embed.dim, dev=dev)
        L.time_sinusoid(t_vec.contiguous(), B, dec0.time_embed.dim, dec0.time_embed.freqs(dev), 1000.0, emb)
        lens32 = None if lens is None else lens.to(torch.int32).contiguous()
        brs = []
        for bw in pk.branches:
            d = bw.dec
            C, nl = bw.C, bw.nl
            w = type("S", (), {})()
            w.F = 1 + T // bw.hop
            w.R = R = B * w.F
            w.mask = None
            if lens32 is not None:
                w.mask = _e(R, dev=dev)
                L.frame_mask(lens32, B, w.F, bw.hop, w.mask)
            # ---- conditioning path at mel rate (cond_mlp + all cond_proj) -----------------
            cm = d.cond_mlp
            w.cm_hpre, w.cm_h = _e(Rc, bw.ch, dev=dev), _e(Rc, bw.ch, dev=dev)
            L.gemm_group([_nt(c0, bw.cc, bw.cmW0, bw.cc, w.cm_h, bw.ch, Rc, bw.ch, bw.cc, bn=BN_G1,
                              bias=cm[0].bias.data_ptr(), slope=cm[1].weight.data_ptr(), act=L.ACT_PRELU,
                              round_tf32=1, c_pre=w.cm_hpre.data_ptr(), ld_pre=bw.ch)])
            w.c1 = _e(Rc, bw.cc, dev=dev)
            L.gemm_group([_nt(w.cm_h, bw.ch, bw.cmW2, bw.ch, w.c1, bw.cc, Rc, bw.cc, bw.ch, bn=BN_G2,
                              bias=cm[2].bias.data_ptr(), round_tf32=1)])
            w.cp = _e(Rc, nl * C, dev=dev)
            L.gemm_group([_nt(w.c1, bw.cc, bw.Wcp, bw.cc, w.cp, nl * C, Rc, nl * C, bw.cc, bn=BN_G1,
                              bias=bw.bcp.data_ptr())])
            # ---- time path ------------------------------------------------------------------
            tm = d.time_mlp
            Ht = tm[0].weight.shape[0]
            w.te1pre, w.te1 = _e(B, Ht, dev=dev), _e(B, Ht, dev=dev)
            L.linear_small(emb, B, bw.te, bw.te, tm[0].weight, bw.te, tm[0].bias, Ht, L.ACT_NONE, w.te1pre, Ht)
            L.linear_small(emb, B, bw.te, bw.te, tm[0].weight, bw.te, tm[0].bias, Ht, L.ACT_SILU, w.te1, Ht)
            w.te2 = _e(B, bw.te, dev=dev)
            L.linear_small(w.te1, B, Ht, Ht, tm[2].weight, Ht, tm[2].bias, bw.te, L.ACT_NONE, w.te2, bw.te)
            w.ts = _e(B, nl * C, dev=dev)
            L.linear_small(w.te2, B, bw.te, bw.te, bw.Wte, bw.te, bw.bte, nl * C, L.ACT_NONE, w.ts, nl * C)
            # ---- STFT -> in_proj -> in_norm -> blocks -> out_proj -> irfft frames -----------
            w.pin = _e(R, bw.ldp, dev=dev)
            L.stft(x_audio, B, T, T, bw.n_fft, bw.hop, L.SPEC_PACKED, w.pin, bw.ldp, round_tf32=1)
            w.x0 = _e(R, C, dev=dev)
            L.gemm_group([_nt(w.pin, bw.ldp, bw.Win, bw.ldp, w.x0, C, R, C, bw.nin, bias=d.in_proj.bias.data_ptr())])
            xs = _e(R, C, dev=dev)
            w.inv0 = _e(R, dev=dev)
            w.lim0 = _limit_draw(training)
            L.biasnorm(w.x0, R, C, C, d.in_norm.bias, d.in_norm.log_scale, xs, C, w.inv0)
            w.xl, w.saves = _blocks_forward(bw.blocks, xs, B, w.F, C, w.mask, w.cp, nl * C, Fm, bw.factor,
                                            B * Fm, w.ts, training, round_last=True)
            pout = _e(R, bw.ldp, dev=dev)
            L.gemm_group([_nt(w.xl, C, bw.Wout, C, pout, bw.ldp, R, bw.nin, C, bias=d.out_proj.bias.data_ptr(),
                              row_scale=L.ptr(w.mask))])
            w.fr = _e(R, bw.n_fft, dev=dev)
            L.irfft_frames(pout, R, bw.ldp, bw.n_fft, w.fr)
            brs.append(w)
        pred = _e(B, T, dev=dev)
        wgt = None if weight is None else weight.contiguous()
        L.ola_combine([w.fr for w in brs], [bw.n_fft for bw in pk.branches], [bw.hop for bw in pk.branches],
                      [w.F for w in brs], wgt, None, pred, B, T, False, 0.0, 0.0, False)
        for w in brs:
            w.fr = None
        ctx.model = model
        ctx.saved = (c0, brs, emb, wgt, B, T, Fm)
        return pred

    @staticmethod
    def backward(ctx, dpred: Tensor):
        with _zero_arena(dpred.device):
            return _ProcessModelFn._backward(ctx, dpred)

    @staticmethod
    def _backward(ctx, dpred: Tensor):
        model = ctx.model
        pk: PackedGenerator = model._packed
        c0, brs, emb, wgt, B, T, Fm = ctx.saved
        dev = dpred.device
        dpred = dpred.contiguous()
        Rc = c0.shape[0]
        grads: Dict[str, Tensor] = {}
        dc0 = _z(Rc, c0.shape[1], dev=dev)
        dx_audio = _z(B, T, dev=dev) if ctx.needs_input_grad[2] else None
        nb = len(pk.branches)
        for j, (bw, w) in enumerate(zip(pk.branches, brs)):
            d = bw.dec
            C, nl, R, F = bw.C, bw.nl, w.R, w.F
            pre = f"estimators.{j}.decoder."
            # ---- iSTFT + branch fusion adjoint ----------------------------------------------
            if wgt is None:
                g = dpred
                scale = 1.0 / nb
            else:
                g = (dpred * wgt[:, j:j + 1]).contiguous()
                scale = 1.0
            Lp = bw.n_fft + bw.hop * (F - 1)
            gsig = _e(B, Lp, dev=dev)
            L.istft_bwd_prep(g, B, T, bw.n_fft, bw.hop, F, scale, gsig)
            dpout = _e(R, bw.ldp, dev=dev)
            L.istft_bwd_spec(gsig, B, Lp, bw.n_fft, bw.hop, w.mask, dpout, bw.ldp, round_tf32=1)
            g_bo = _z(bw.nin, dev=dev)
            L.colsum(dpout, bw.ldp, R, bw.nin, g_bo)
            gWo = _e(bw.nin, C, dev=dev)
            L.gemm_group([_tn(dpout, bw.ldp, w.xl, C, gWo, C, bw.nin, C, R)])
            dxl = _e(R, C, dev=dev)
            L.gemm_group([_nn(dpout, bw.ldp, bw.Wout, C, dxl, C, R, C, bw.nin, bn=BN_G2)])
            grads[pre + "out_proj.weight"] = gWo.view(bw.nin, C, 1)
            grads[pre + "out_proj.bias"] = g_bo
            # ---- blocks -----------------------------------------------------------------------
            g_ts = _z(B, nl * C, dev=dev)
            dcp = _z(Rc, nl * C, dev=dev)
            dxs = _blocks_backward(bw.blocks, w.saves, dxl, B, F, C, w.mask, w.cp, nl * C, Fm, bw.factor,
                                   B * Fm, w.ts, g_ts, dcp, grads, pre)
            dx0 = _norm_backward(dxs, w.x0, w.inv0, d.in_norm, R, C, B, F, w.lim0, grads, pre + "in_norm.")
            g_bi = _z(C, dev=dev)
            L.colsum(dx0, C, R, C, g_bi)
            gWi = _e(C, bw.nin, dev=dev)
            L.gemm_group([_tn(dx0, C, w.pin, bw.ldp, gWi, bw.nin, C, bw.nin, R)])
            grads[pre + "in_proj.weight"] = gWi.view(C, bw.nin, 1)
            grads[pre + "in_proj.bias"] = g_bi
            if dx_audio is not None:
                dpin = _e(R, bw.ldp, dev=dev)
                L.gemm_group([_nn(dx0, C, bw.Win, bw.ldp, dpin, bw.ldp, R, bw.nin, C)])
                frg = _e(R, bw.n_fft, dev=dev)
                L.stft_bwd_frames(dpin, R, bw.ldp, bw.n_fft, frg)
                L.stft_bwd_fold(frg, B, T, bw.n_fft, bw.hop, F, dx_audio, True)
            # ---- time path ------------------------------------------------------------------
            tm = d.time_mlp
            Ht, te = tm[0].weight.shape[0], bw.te
            gWte = _e(nl * C, te, dev=dev)
            L.gemm_group([_tn(g_ts, nl * C, w.te2, te, gWte, te, nl * C, te, B)])
            g_bte = _z(nl * C, dev=dev)
            L.colsum(g_ts, nl * C, B, nl * C, g_bte)
            dte2 = _e(B, te, dev=dev)
            L.gemm_group([_nn(g_ts, nl * C, bw.Wte, te, dte2, te, B, te, nl * C)])
            gW2t = _e(te, Ht, dev=dev)
            L.gemm_group([_tn(dte2, te, w.te1, Ht, gW2t, Ht, te, Ht, B)])
            g_b2t = _z(te, dev=dev)
            L.colsum(dte2, te, B, te, g_b2t)
            dte1 = _e(B, Ht, dev=dev)
            L.gemm_group([_nn(dte2, te, tm[2].weight, Ht, dte1, Ht, B, Ht, te)])
            g_b0t = _z(Ht, dev=dev)
            L.act_bwd(dte1, Ht, w.te1pre, Ht, None, 0.0, L.ACT_SILU, B, Ht, dte1, Ht, g_b0t, None)
            gW0t = _e(Ht, te, dev=dev)
            L.gemm_group([_tn(dte1, Ht, emb, te, gW0t, te, Ht, te, B)])
            grads[pre + "time_mlp.0.weight"] = gW0t
            grads[pre + "time_mlp.0.bias"] = g_b0t
            grads[pre + "time_mlp.2.weight"] = gW2t
            grads[pre + "time_mlp.2.bias"] = g_b2t
            for i in range(nl):
                grads[f"{pre}blocks.{i}.time_embed_proj.weight"] = gWte[i * C:(i + 1) * C]
                grads[f"{pre}blocks.{i}.time_embed_proj.bias"] = g_bte[i * C:(i + 1) * C]
            # ---- conditioning path ------------------------------------------------------------
            cm = d.cond_mlp
            cc, ch = bw.cc, bw.ch
            g_bcp = _z(nl * C, dev=dev)
            L.colsum(dcp, nl * C, Rc, nl * C, g_bcp)
            gWcp = _e(nl * C, cc, dev=dev)
            L.gemm_group([_tn(dcp, nl * C, w.c1, cc, gWcp, cc, nl * C, cc, Rc)])
            dc1 = _e(Rc, cc, dev=dev)
            L.gemm_group([_nn(dcp, nl * C, bw.Wcp, cc, dc1, cc, Rc, cc, nl * C)])
            for i in range(nl):
                grads[f"{pre}blocks.{i}.cond_proj.weight"] = gWcp[i * C:(i + 1) * C].view(C, cc, 1)
                grads[f"{pre}blocks.{i}.cond_proj.bias"] = g_bcp[i * C:(i + 1) * C]
            g_b2c = _z(cc, dev=dev)
            L.colsum(dc1, cc, Rc, cc, g_b2c)
            gWc2 = _e(cc, ch, dev=dev)
            L.gemm_group([_tn(dc1, cc, w.cm_h, ch, gWc2, ch, cc, ch, Rc)])
            dcmh = _e(Rc, ch, dev=dev)
            L.gemm_group([_nn(dc1, cc, bw.cmW2, ch, dcmh, ch, Rc, ch, cc, bn=BN_G1)])
            g_b0c, g_slc = _z(ch, dev=dev), _z(ch, dev=dev)
            L.act_bwd(dcmh, ch, w.cm_hpre, ch, cm[1].weight, 0.0, L.ACT_PRELU, Rc, ch, dcmh, ch, g_b0c, g_slc,
                      round_tf32=1)
            gWc0 = _e(ch, cc, dev=dev)
            L.gemm_group([_tn(dcmh, ch, c0, cc, gWc0, cc, ch, cc, Rc)])
            L.gemm_group([_nn(dcmh, ch, bw.cmW0, cc, dc0, cc, Rc, cc, ch, accumulate=1)])
            grads[pre + "cond_mlp.0.weight"] = gWc0.view(ch, cc, 1)
            grads[pre + "cond_mlp.0.bias"] = g_b0c
            grads[pre + "cond_mlp.1.weight"] = g_slc
            grads[pre + "cond_mlp.2.weight"] = gWc2.view(cc, ch, 1)
            grads[pre + "cond_mlp.2.bias"] = g_b2c
        names = model._est_param_names
        return (None, dc0, dx_audio, None, None, None) + tuple(grads.get(n) for n in names)


# =========================================================================================
# public entry points used by generator.py
# =========================================================================================
def _param_lists(model):
    if not hasattr(model, "_ce_param_names"):
        model._ce_param_names = [n for n, _ in model.named_parameters() if n.startswith("cond_encoder.")]
        model._est_param_names = [n for n, _ in model.named_parameters() if n.startswith("estimators.")]
    pd = dict(model.named_parameters())
    return [pd[n] for n in model._ce_param_names], [pd[n] for n in model._est_param_names]


def encode_cond(model, mel: Tensor) -> Tensor:
    """(B, n_mels, F) mel -> (B*F + 1, C) conditioning rows (last row all-zero), differentiable."""
    ce_p, _ = _param_lists(model)
    return _CondEncoderFn.apply(model, mel.float(), *ce_p)


def process_model(model, c0: Tensor, x: Tensor, t: Tensor, audio_lens: Optional[Tensor],
                  weight: Optional[Tensor]) -> Tensor:
    _, est_p = _param_lists(model)
    return _ProcessModelFn.apply(model, c0, x.float(), t.reshape(-1).float(), audio_lens, weight, *est_p)


def _branch_dropout_weight(model, batch_size: int, device) -> Optional[Tensor]:
    """generator.py:145-162 (same draws, same order), as per-sample branch weights incl. the mean."""
    nb = model.num_branches
    if not (model.training and model.branch_dropout > 0.0 and nb > 1):
        return None
    branch_idx = torch.randint(0, nb, (batch_size,), device=device)
    mask = torch.ones((batch_size, nb), device=device)
    mask[torch.arange(batch_size, device=device), branch_idx] = 0.0
    mask = mask * (nb / (nb - 1))
    weight = torch.where(torch.rand((batch_size, 1), device=device) < model.branch_dropout, mask,
                         torch.ones_like(mask))
    return weight / nb


def generator_infer_with_grad(model, mel: Tensor, noise: Tensor, audio_lens: Optional[Tensor],
                              n_timesteps: int, clamp_pred: bool) -> Tensor:
    """Euler sampler with autograd (generator.py:236-271); the Euler update itself is three
    element-wise torch ops per step on (B, T)."""
    c0 = encode_cond(model, mel)
    t_span = torch.linspace(0, 1, n_timesteps + 1, device=noise.device)
    t, dt = t_span[0], t_span[1] - t_span[0]
    x = noise.float()
    B = noise.shape[0]
    for step in range(1, n_timesteps + 1):
        w = _branch_dropout_weight(model, B, noise.device)
        pred = process_model(model, c0, x, t.expand(B), audio_lens, w)
        x = x + (pred - x) / (1 - t) * dt
        t = t_span[step]
    return x.clamp(min=-1.0, max=1.0) if clamp_pred else x


def generator_fm_loss(model, cond: Tensor, audio: Tensor, audio_lens: Tensor,
                      noise: Optional[Tensor] = None, t: Optional[Tensor] = None) -> Tensor:
    """MelAudioGenerator.forward: flow-matching endpoint loss with spectral-energy scaling
    (generator.py:294-325, 202-234, 172-200).  `noise` / `t` (extensions) pin the random draws."""
    from .losses import spectral_scaled_loss
    model._require_cuda()
    cond = model._maybe_noisy_cond(cond)
    c0 = encode_cond(model, cond)
    if noise is None:
        noise = torch.randn_like(audio) * model.init_noise_scale
    if t is None:
        t = torch.rand((audio.shape[0], 1), device=audio.device, dtype=audio.dtype)
    t = t.reshape(-1, 1)
    x = (1.0 - t) * noise + t * audio
    w = _branch_dropout_weight(model, audio.shape[0], audio.device)
    pred = process_model(model, c0, x, t, audio_lens, w)
    return spectral_scaled_loss(model, pred, audio, audio_lens)
