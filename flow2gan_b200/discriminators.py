"""Multi-Period and Multi-Resolution discriminators with the reference's module tree / parameter
names (flow2gan/models/discriminators.py:18-219; weight_norm disabled there, :13-15), computed on
channel-last tensors: the wide convs are windowed ("implicit im2col") tcgen05 TF32 GEMMs (convwin.py),
the first layer of every stack (Cin = 1 / 2) is a direct fp32 convolution (csrc/conv.cu::conv_small_*);
the gather (im2col + GEMM + col2im) path remains as the general fallback and A/B reference."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from . import convwin

# F2G_CONV_IM2COL=1 forces the gather (im2col) path everywhere -- A/B testing only
_USE_WINDOWED = os.environ.get("F2G_CONV_IM2COL", "0") != "1"


def _ceil4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Conv2dCLFn(torch.autograd.Function):
    """y = leaky_relu(conv2d(x) + b) on channel-last (Nb, H, W, C) -> (Nb, Ho, Wo, Co)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, sh: int, sw: int, ph: int, pw: int,
                leaky: Optional[float], packs: Optional[dict] = None):
        Nb, H, W, Cc = x.shape
        Co, Ci, kh, kw = weight.shape
        assert Ci == Cc and x.stride(3) == 1 and x.stride(2) == Cc, (x.shape, x.stride())
        dev = x.device
        K = kh * kw * Cc
        ldk, Cop = _ceil4(K), _ceil4(Co)
        geom = L.conv_geom(Nb, H, W, Cc, x.stride(1), x.stride(0), kh, kw, sh, sw, ph, pw, ldk)
        Ho = (H + 2 * ph - kh) // sh + 1
        Wo = (W + 2 * pw - kw) // sw + 1
        M = Nb * Ho * Wo
        Wp = packs.get("wp") if packs is not None else None       # per-module cache, see convwin._ConvWinFn
        if Wp is None or packs.get("wp_ver") != weight._version:
            if Wp is None:
                Wp = torch.empty(Cop, ldk, device=dev)
            L.conv_w_pack(weight.detach().contiguous(), Co, Ci, kh * kw, Cop, ldk, Wp, 0)
            if packs is not None:
                packs["wp"], packs["wp_ver"] = Wp, weight._version
        col = torch.empty(M, ldk, device=dev)
        L.im2col2d(x.data_ptr(), geom, col, 1)
        y = torch.zeros(M, Cop, device=dev) if Cop != Co else torch.empty(M, Cop, device=dev)
        bn = 64 if Co <= 64 else (128 if Co <= 128 else 256)
        L.gemm_group([L.gemm_desc(col.data_ptr(), Wp.data_ptr(), y.data_ptr(), M, Co, K, ldk, ldk, Cop, bn=bn,
                                  bias=bias.data_ptr(), act=L.ACT_LEAKY if leaky is not None else L.ACT_NONE,
                                  leaky=leaky or 0.0)])
        ctx.saved = (col, y, Wp, geom, (Nb, H, W, Cc, Co, Ci, kh, kw, Ho, Wo, M, K, ldk, Cop), leaky)
        return y.view(Nb, Ho, Wo, Cop)[..., :Co]

    @staticmethod
    def backward(ctx, dy: Tensor):
        col, y, Wp, geom, dims, leaky = ctx.saved
        Nb, H, W, Cc, Co, Ci, kh, kw, Ho, Wo, M, K, ldk, Cop = dims
        dev = dy.device
        dz = torch.zeros(M, Cop, device=dev)
        dz[:, :Co].copy_(dy.reshape(M, Co))
        g_bias = torch.zeros(Cop, device=dev)
        L.act_bwd(dz, Cop, y if leaky is not None else None, Cop, None, leaky or 0.0,
                  L.ACT_LEAKY if leaky is not None else L.ACT_NONE, M, Co, dz, Cop, g_bias, None, round_tf32=1)
        gW = gb = gx = None
        if ctx.needs_input_grad[1]:
            dWp = torch.zeros(Cop, ldk, device=dev)
            split = L.pick_split_k(Co, K, M)
            L.gemm_group([L.gemm_desc(dz.data_ptr(), col.data_ptr(), dWp.data_ptr(), Co, K, M, Cop, ldk, ldk,
                                      bn=128, a_mn=1, b_mn=1, split_k=split)])
            gW = torch.empty(Co, Ci, kh, kw, device=dev)
            L.conv_w_pack(dWp, Co, Ci, kh * kw, Cop, ldk, gW, 1)
        if ctx.needs_input_grad[2]:
            gb = g_bias[:Co]
        if ctx.needs_input_grad[0]:
            dcol = torch.empty(M, ldk, device=dev)
            bn = 64 if K <= 64 else (128 if K <= 128 else 256)
            L.gemm_group([L.gemm_desc(dz.data_ptr(), Wp.data_ptr(), dcol.data_ptr(), M, K, Co, Cop, ldk, ldk,
                                      bn=bn, b_mn=1)])
            if ldk != K:
                dcol[:, K:].zero_()
            gx = torch.empty(Nb, H, W, Cc, device=dev)
            g2 = L.conv_geom(Nb, H, W, Cc, W * Cc, H * W * Cc, kh, kw, geom.sh, geom.sw, geom.ph, geom.pw, ldk)
            L.col2im2d(dcol, g2, gx.data_ptr(), 0)
        return gx, gW, gb, None, None, None, None, None, None


class _ConvSmallFn(torch.autograd.Function):
    """First layer of a discriminator stack (Cin = 1 or 2 -> 32 channels, LeakyReLU): direct fp32
    convolution (csrc/conv.cu::conv_small_*), no im2col matrix, no TF32 rounding.  x: channel-last
    (Nb, H, W, Cin) view with unit channel stride; returns (Nb, Ho, Wo, 32)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, sh: int, sw: int, ph: int, pw: int,
                leaky: Optional[float]):
        Nb, H, W, Cin = x.shape
        Co, _, kh, kw = weight.shape
        assert x.stride(3) == 1 and x.dtype == torch.float32
        w = weight.detach().contiguous()
        Ho = (H + 2 * ph - kh) // sh + 1
        Wo = (W + 2 * pw - kw) // sw + 1
        y = torch.empty(Nb * Ho * Wo, Co, device=x.device, dtype=torch.float32)
        pitches = (x.stride(0), x.stride(1), x.stride(2))
        L.conv_small_fwd(x, Nb, H, W, Cin, pitches, w, bias.detach(), Co, kh, kw, sh, sw, ph, pw, leaky, y)
        ctx.saved = (x.detach(), w, y, (Nb, H, W, Cin, Co, kh, kw, sh, sw, ph, pw, Ho, Wo), pitches, leaky)
        return y.view(Nb, Ho, Wo, Co)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w, y, dims, pitches, leaky = ctx.saved
        Nb, H, W, Cin, Co, kh, kw, sh, sw, ph, pw, Ho, Wo = dims
        dy = dy.contiguous()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        K = kh * kw * Cin
        acc = torch.zeros(K * Co + Co, device=dy.device, dtype=torch.float32) if need_w else None
        gwp = acc[:K * Co] if need_w else None
        gbv = acc[K * Co:] if need_w else None
        dx = torch.empty(Nb, H, W, Cin, device=dy.device, dtype=torch.float32) if need_x else None
        scratch = None
        if need_w:       # one partial record per block of the weight-gradient kernel (deterministic reduction)
            tiles = Nb * Ho * ((Wo + 31) // 32)
            scratch = torch.empty(min(tiles, 148 * 8) * (K * Co + Co), device=dy.device, dtype=torch.float32)
        L.conv_small_bwd(x, Nb, H, W, Cin, pitches, w, Co, kh, kw, sh, sw, ph, pw, leaky, dy, y, gwp, gbv, dx, scratch)
        gW = gwp.view(kh, kw, Cin, Co).permute(3, 2, 0, 1).contiguous() if ctx.needs_input_grad[1] else None
        gb = gbv if ctx.needs_input_grad[2] else None
        return dx, gW, gb, None, None, None, None, None


def _small_ok(x: Tensor, w: Tensor, sw: int = 1) -> bool:
    return (x.shape[3] in (1, 2) and w.shape[0] == 32 and w.shape[2] * w.shape[3] * x.shape[3] <= 56
            and w.shape[2] <= 3 and w.shape[3] <= 9 and sw <= 3 and x.stride(3) == 1 and x.dtype == torch.float32)


def _pack_cache(conv: nn.Conv2d, weight: Tensor, key: str) -> dict:
    """GEMM-ready weight layouts of one Conv2d, kept on the module and rebuilt IN PLACE only when the
    parameter's version (optimizer steps bump it) or storage changed: a D+G iteration pair runs every
    discriminator conv three times (real|fake in the D phase, real and fake in the G phase) but its
    weights change once -- packing per call was 4.5 % of the pair.  In place: CUDA graphs that captured
    the buffers keep reading current data whichever graph (or eager call) refreshed them last."""
    store = conv.__dict__.setdefault("_f2g_packs", {})
    ent = store.get(key)
    if ent is None or ent.get("ptr") != weight.data_ptr():
        ent = store[key] = {"ptr": weight.data_ptr()}
    return ent


def conv2d_cl(x: Tensor, conv: nn.Conv2d, leaky: Optional[float], train_weights: bool = True,
              swap_hw: bool = False) -> Tensor:
    """Channel-last Conv2d (+LeakyReLU).  `swap_hw`: x is laid out (Nb, W, H, C) -- the module's
    kernel / stride / padding pairs are applied with H and W exchanged (DiscriminatorP runs its
    (k, 1) convs along the contiguous axis of a (B*period, 1, T/period, C) tensor).  Convs the
    windowed path supports (C % 32 == 0, unit H stride) never materialise im2col (convwin.py)."""
    w, b = conv.weight, conv.bias
    if not train_weights:
        w, b = w.detach(), b.detach()
    sh, sw = conv.stride
    ph, pw = conv.padding
    if swap_hw:
        w = w.transpose(2, 3)
        sh, sw, ph, pw = sw, sh, pw, ph
    if _USE_WINDOWED and convwin.supports(x.shape[3], w.shape[2], w.shape[3], sh, sw):
        return convwin.conv2d_win(x, w, b, sw, ph, pw, leaky, _pack_cache(conv, w, "win%d" % swap_hw))
    if _USE_WINDOWED and _small_ok(x, w, sw):      # first layers (Cin = 1 / 2): direct convolution, no im2col
        return _ConvSmallFn.apply(x, w, b, sh, sw, ph, pw, leaky)
    return _Conv2dCLFn.apply(x, w, b, sh, sw, ph, pw, leaky, _pack_cache(conv, w, "col%d" % swap_hw))


class DiscriminatorP(nn.Module):
    def __init__(self, period: int, in_channels: int = 1, kernel_size: int = 5, stride: int = 3,
                 lrelu_slope: float = 0.1, num_embeddings: Optional[int] = None):
        super().__init__()
        if num_embeddings is not None:
            raise NotImplementedError("conditional discriminators are not used by the GAN recipe")
        self.period = period
        k, s = kernel_size, stride
        self.convs = nn.ModuleList([
            nn.Conv2d(in_channels, 32, (k, 1), (s, 1), padding=(k // 2, 0)),
            nn.Conv2d(32, 128, (k, 1), (s, 1), padding=(k // 2, 0)),
            nn.Conv2d(128, 512, (k, 1), (s, 1), padding=(k // 2, 0)),
            nn.Conv2d(512, 1024, (k, 1), (s, 1), padding=(k // 2, 0)),
            nn.Conv2d(1024, 1024, (k, 1), (1, 1), padding=(k // 2, 0)),
        ])
        self.conv_post = nn.Conv2d(1024, 1, (3, 1), 1, padding=(1, 0))
        self.lrelu_slope = lrelu_slope

    def forward(self, x: Tensor, train_weights: bool = True) -> Tuple[Tensor, List[Tensor]]:
        """x (B, T) -> (score (B, -1), fmap list); feature maps are channel-last (B, H, p, C) views.
        Internally the tensor is kept period-major, (B*p, 1, T/p, C): the (k, 1) convs then slide
        along the contiguous axis and take the windowed (no im2col) path."""
        b, t = x.shape
        p = self.period
        if t % p != 0:
            n_pad = p - (t % p)
            x = torch.nn.functional.pad(x.unsqueeze(1), (0, n_pad), "reflect").squeeze(1)
            t += n_pad
        h = x.reshape(b, t // p, p).transpose(1, 2).reshape(b * p, 1, t // p, 1)

        def as_bhpc(v: Tensor) -> Tensor:
            return v.unflatten(0, (b, p)).squeeze(2).permute(0, 2, 1, 3)

        fmap = []
        for i, l in enumerate(self.convs):
            h = conv2d_cl(h, l, self.lrelu_slope, train_weights, swap_hw=True)
            if i > 0:
                fmap.append(as_bhpc(h))
        h = conv2d_cl(h, self.conv_post, None, train_weights, swap_hw=True)
        h = as_bhpc(h)
        fmap.append(h)
        return h.reshape(b, -1), fmap


class MultiPeriodDiscriminator(nn.Module):
    def __init__(self, periods: Tuple[int, ...] = (2, 3, 5, 7, 11), num_embeddings: Optional[int] = None):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorP(period=p, num_embeddings=num_embeddings)
                                             for p in periods])

    def forward(self, y: Tensor, y_hat: Tensor, bandwidth_id=None, train_weights: bool = True):
        return _run_real_fake(self.discriminators, y, y_hat, train_weights)


def _run_real_fake(discs, y: Tensor, y_hat: Tensor, train_weights: bool):
    """D-phase (weights trained): one pass over the concatenated real|fake batch.  G-phase
    (weights frozen, gradient only w.r.t. y_hat): the real half runs without autograd so no
    input-gradient work is spent on it (the reference runs d(y), d(y_hat) separately too)."""
    if train_weights or not y_hat.requires_grad:
        outs = [d(torch.cat([y, y_hat], 0), train_weights) for d in discs]
        return _split_real_fake(outs, y.shape[0])
    with torch.no_grad():
        real = [d(y, False) for d in discs]
    fake = [d(y_hat, False) for d in discs]
    return ([s for s, _ in real], [s for s, _ in fake], [f for _, f in real], [f for _, f in fake])


def _split_real_fake(outs, B):
    y_d_rs = [s[:B] for s, _ in outs]
    y_d_gs = [s[B:] for s, _ in outs]
    fmap_rs = [[f[:B] for f in fm] for _, fm in outs]
    fmap_gs = [[f[B:] for f in fm] for _, fm in outs]
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


class _ComplexSpecFn(torch.autograd.Function):
    """(B, T) -> (B, frames, freq, 2) complex STFT (hop = n/4), channel-last, with the STFT adjoint."""

    @staticmethod
    def forward(ctx, x: Tensor, n_fft: int, hop: int):
        x = x.contiguous().float()
        B, T = x.shape
        frames, nb = 1 + T // hop, n_fft // 2 + 1
        out = torch.empty(B * frames, 2 * nb, device=x.device)
        L.stft(x, B, T, T, n_fft, hop, L.SPEC_COMPLEX, out, 2 * nb)
        ctx.cfg = (B, T, n_fft, hop, frames, nb)
        return out.view(B, frames, nb, 2)

    @staticmethod
    def backward(ctx, d: Tensor):
        B, T, n_fft, hop, frames, nb = ctx.cfg
        d = d.contiguous().view(B * frames, 2 * nb)
        fr = torch.empty(B * frames, n_fft, device=d.device)
        L.stft_bwd_frames(d, B * frames, 2 * nb, n_fft, fr, interleaved=1)
        dx = torch.empty(B, T, device=d.device)
        L.stft_bwd_fold(fr, B, T, n_fft, hop, frames, dx, False)
        return dx, None, None


class _SpecHolder(nn.Module):
    def __init__(self, n_fft: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class DiscriminatorR(nn.Module):
    def __init__(self, window_length: int, num_embeddings: Optional[int] = None, channels: int = 32,
                 hop_factor: float = 0.25,
                 bands=((0.0, 0.1), (0.1, 0.25), (0.25, 0.5), (0.5, 0.75), (0.75, 1.0))):
        super().__init__()
        if num_embeddings is not None:
            raise NotImplementedError("conditional discriminators are not used by the GAN recipe")
        self.window_length = window_length
        self.hop_length = int(window_length * hop_factor)
        self.spec_fn = _SpecHolder(window_length)
        n_bins = window_length // 2 + 1
        self.bands = [(int(b[0] * n_bins), int(b[1] * n_bins)) for b in bands]
        ch = channels

        def convs():
            return nn.ModuleList([
                nn.Conv2d(2, ch, (3, 9), (1, 1), padding=(1, 4)),
                nn.Conv2d(ch, ch, (3, 9), (1, 2), padding=(1, 4)),
                nn.Conv2d(ch, ch, (3, 9), (1, 2), padding=(1, 4)),
                nn.Conv2d(ch, ch, (3, 9), (1, 2), padding=(1, 4)),
                nn.Conv2d(ch, ch, (3, 3), (1, 1), padding=(1, 1)),
            ])
        self.band_convs = nn.ModuleList([convs() for _ in self.bands])
        self.conv_post = nn.Conv2d(ch, 1, (3, 3), (1, 1), padding=(1, 1))

    def spectrogram(self, x: Tensor) -> Tensor:
        # DC removal + peak normalisation (discriminators.py:187-190): two reductions over (B, T)
        x = x - x.mean(dim=-1, keepdim=True)
        x = 0.8 * x / (x.abs().max(dim=-1, keepdim=True)[0] + 1e-9)
        return _ComplexSpecFn.apply(x, self.window_length, self.hop_length)      # (B, frames, freq, 2)

    def forward(self, x: Tensor, train_weights: bool = True):
        spec = self.spectrogram(x)
        fmap, outs = [], []
        for (lo, hi), stack in zip(self.bands, self.band_convs):
            band = spec[:, :, lo:hi, :]
            for i, layer in enumerate(stack):
                band = conv2d_cl(band, layer, 0.1, train_weights)
                if i > 0:
                    fmap.append(band)
            outs.append(band)
        h = torch.cat(outs, dim=2)
        h = conv2d_cl(h, self.conv_post, None, train_weights)
        fmap.append(h)
        return h, fmap


class MultiResolutionDiscriminator(nn.Module):
    def __init__(self, fft_sizes: Tuple[int, ...] = (2048, 1024, 512), num_embeddings: Optional[int] = None):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorR(window_length=w, num_embeddings=num_embeddings)
                                             for w in fft_sizes])

    def forward(self, y: Tensor, y_hat: Tensor, bandwidth_id=None, train_weights: bool = True):
        return _run_real_fake(self.discriminators, y, y_hat, train_weights)
