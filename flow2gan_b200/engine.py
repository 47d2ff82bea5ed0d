"""Inference engine: runs MelAudioGenerator.infer / BaseAudioGenerator.infer
(flow2gan/models/generator.py:236-271,327-366) as a fixed launch sequence of the CUDA library
(csrc/*.cu) over pre-packed weights and pre-allocated channel-last workspaces, captured in a
CUDA graph per (batch, frames, length, n_timesteps) shape.

What is restructured relative to the reference (results stay equal):
  * activations are channel-last rows (b*T + t, C) so 1x1 convs are K-major TF32 GEMMs;
  * cond_mlp / cond_proj (step-invariant, and computed by the reference at the upsampled frame
    rate every ODE step) run once per call at mel-frame rate -- they commute with the frame
    repetition of upsample_cond (modules.py:668-680); the zero-padded tail frame is served by
    one extra all-zero conditioning row;
  * the three branches' same-stage GEMMs are one grouped persistent launch.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L


import operator
import os

BN_G1 = 128   # N-tile hint of pwconv1-like GEMMs (wide N); the CTA-pair kernel picks the byte-optimal tile
BN_G2 = 128   # N-tile hint of pwconv2-like GEMMs (N = channels) unless the problem is a chained consumer

# Operand type of the ConvNeXt-block contractions (pwconv1 / pwconv2) at inference.
#   "f16"  (default): fp16 operands, fp32 accumulate (tcgen05 kind::f16).  fp16 has the SAME 11-bit
#          significand as TF32, so inside the fp16 range the rounding error is the TF32 one
#          (parity tests: same rel-RMS), while the operand bytes through the L2->SM fabric --
#          the measured bound of these short-K GEMMs -- halve and the MMA rate doubles.
#          Values are clamped to +-65504 where they are produced.
#   "tf32": fp32-container TF32 operands everywhere (8-bit exponent; the training path).
BLOCK_OPERANDS = os.environ.get("F2G_BLOCK_OPERANDS", "f16")
assert BLOCK_OPERANDS in ("f16", "tf32"), BLOCK_OPERANDS
# pwconv1 -> pwconv2 of a block as ONE chained launch (per-row-tile counters instead of a kernel
# boundary; fp16 operands only)
CHAIN_MLP = os.environ.get("F2G_CHAIN_MLP", "1") == "1"
FORK_COND = os.environ.get("F2G_FORK_COND", "1") == "1"    # conditioning path on a second stream
# The sampler evaluates every batch element at the same t_k = k / N, so the per-layer time-scale
# vectors of step k depend on the weights only: they are computed once per (plan, N, weight version)
# instead of in every model evaluation (4 small launches per ODE step leave the launch graph; output
# bit-identical, profiles/r02_switches.md).  F2G_CACHE_TIME=0 restores the per-step launches (A/B).
CACHE_TIME = os.environ.get("F2G_CACHE_TIME", "1") == "1"


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


# Chained GEMM launches (pwconv1 -> pwconv2 through per-row-tile counters, csrc/gemm_pair.cu) need all
# CTAs of the launch resident: consumer tiles spin until the producer tiles -- possibly in CTAs that
# are not scheduled yet -- have stored their rows.  Two such launches running at the same time on two
# streams can starve each other of SMs.  Per device, the host therefore serialises every launch
# sequence that contains chained launches (an eager `_run`, a graph replay) across streams: the next
# one waits for an event recorded behind the previous one whenever that was issued on another stream.
_CHAIN_GUARD: Dict[int, Tuple[torch.cuda.Event, int]] = {}


class _chain_guard:
    """with _chain_guard(device): <issue launches containing chained GEMMs on the current stream>"""

    def __init__(self, device: torch.device):
        self.cuda = device.type == "cuda"
        self.idx = device.index if device.index is not None or not self.cuda else torch.cuda.current_device()

    def __enter__(self):
        # inside an outer capture: one stream by construction
        self.skip = (not self.cuda) or torch.cuda.is_current_stream_capturing()
        if self.skip:
            return self
        cur = torch.cuda.current_stream(self.idx)
        prev = _CHAIN_GUARD.get(self.idx)
        if prev is not None and prev[1] != cur.cuda_stream:
            cur.wait_event(prev[0])
        return self

    def __exit__(self, *exc):
        if not self.skip:
            cur = torch.cuda.current_stream(self.idx)
            prev = _CHAIN_GUARD.get(self.idx)
            ev = prev[0] if prev is not None else torch.cuda.Event()
            ev.record(cur)
            _CHAIN_GUARD[self.idx] = (ev, cur.cuda_stream)
        return False


_version_of = operator.attrgetter("_version")


def _params_signature(plist) -> Tuple[int, int]:
    """Cheap change detector over a cached parameter list: in-place updates (optimizer steps,
    load_state_dict) bump tensor versions; storage moves go through Module._apply, which drops
    the packed weights altogether.  (Walking model.parameters() each call cost 0.5 ms.)"""
    return sum(map(_version_of, plist)), plist[0].data_ptr()


def _pack(src: Tensor, rows: int, cols: int, rs: int, cs: int, ld: int, rnd: int = 1,
          dst: Optional[Tensor] = None, dst_col: int = 0, dst_row: int = 0) -> Tensor:
    """dst[dst_row + r, dst_col + c] = src.flat[r*rs + c*cs]  (TF32-rounded by default)."""
    if dst is None:
        dst = torch.zeros(rows, ld, device=src.device, dtype=torch.float32)
    assert src.is_contiguous() and src.dtype == torch.float32
    dptr = dst.data_ptr() + 4 * (dst_row * dst.stride(0) + dst_col)
    L.pack2d(src.data_ptr(), rs, cs, rows, cols, dptr, dst.stride(0), cols, rnd)
    return dst


class _BlockW:
    """Per-ConvNeXtBlock packed matrices + direct pointers to the small vectors."""

    def __init__(self, blk):
        C, H = blk.channels, blk.hidden_channels
        self.C, self.H = C, H
        self.blk = blk
        self.dwT = self.W1 = self.W2 = None
        self._half = None
        self._half_fresh = False
        self.refresh()

    def refresh(self) -> None:
        """(Re)packs IN PLACE: the buffers are allocated once, so every captured CUDA graph that
        holds their addresses (inference plans, the trainer's phase graphs) stays valid."""
        blk, C, H = self.blk, self.C, self.H
        self.dwT = _pack(blk.dwconv.weight.detach(), 7, C, 1, 7, C, rnd=0, dst=self.dwT)     # (7, C)
        self.W1 = _pack(blk.pwconv1.weight.detach(), H, C, C, 1, C, dst=self.W1)             # (H, C)
        self.W2 = _pack(blk.pwconv2.weight.detach(), C, H, H, 1, H, dst=self.W2)             # (C, H)
        self._half_fresh = False
        if self._half is not None:          # already handed out (and maybe baked into a graph): keep current
            self.half()

    def half(self) -> Tuple[Tensor, Tensor]:
        """fp16 copies (RN) of the two matrices for the kind::f16 inference GEMMs: allocated on first
        use (the training path never asks for them), rewritten in place after every refresh."""
        b = self.blk
        if self._half is None:
            dev = b.pwconv1.weight.device
            self._half = (torch.empty(self.H, self.C, device=dev, dtype=torch.float16),
                          torch.empty(self.C, self.H, device=dev, dtype=torch.float16))
            self._half_fresh = False
        if not self._half_fresh:
            self._half[0].copy_(b.pwconv1.weight.detach().reshape(self.H, self.C))
            self._half[1].copy_(b.pwconv2.weight.detach().reshape(self.C, self.H))
            self._half_fresh = True
        return self._half


class PackedGenerator:
    """GEMM-ready (TF32-rounded, 16B-aligned, concatenated) copies of the generator matrices.
    Allocated once; `refresh()` rewrites them IN PLACE when the parameters change (signature over
    tensor versions), so buffer addresses baked into captured CUDA graphs never go stale.
    `version` counts refreshes (plans key their own derived caches on it)."""

    def __init__(self, model):
        self.model = model
        self.signature = None
        self.version = 0
        self._built = False
        self.refresh()

    def stale(self) -> bool:
        return _params_signature(self._plist) != self.signature

    def fp16_weight_error(self) -> float:
        """Largest relative Frobenius error of the fp16 weight copies handed out so far against their
        fp32 sources.  11-bit rounding alone gives ~1.4e-4; overflow (inf) or underflow (weights below
        fp16's 6e-8 subnormal step flushing to zero) shows up as a much larger figure.  One host
        read-back: called once per plan by the fp16 range guard (generator.py), never per step."""
        pairs = []
        for bw in self.ce_blocks + [b for br in self.branches for b in br.blocks]:
            if bw._half is not None:
                blk = bw.blk
                pairs += [(bw._half[0], blk.pwconv1.weight.detach().reshape(bw.H, bw.C)),
                          (bw._half[1], blk.pwconv2.weight.detach().reshape(bw.C, bw.H))]
        for br in self.branches:
            pairs += [(t, getattr(br, nm)) for nm, t in br._halves.items()]
        if not pairs:
            return 0.0
        errs = torch.stack([(h.float() - w).norm() / w.norm().clamp_min(1e-30) for h, w in pairs])
        errs = torch.nan_to_num(errs, nan=float("inf"))
        return float(errs.max())

    def refresh(self) -> None:
        m = self.model
        first = not self._built
        self._plist = list(m.parameters())
        with torch.no_grad():
            ce = m.cond_encoder
            nm = ce.cond_dim
            if first:
                self.ld_mel = _ceil(3 * nm, 4)
                self.ce_Win = torch.zeros(ce.channels, self.ld_mel, device=ce.in_proj.weight.device)
            for k in range(3):   # (Co, Ci, 3) -> column k*Ci + ci
                w = ce.in_proj.weight.detach()
                L.pack2d(w.data_ptr() + 4 * k, 3 * nm, 3, ce.channels, nm,
                         self.ce_Win.data_ptr() + 4 * k * nm, self.ld_mel, nm, 1)
            if first:
                self.ce_blocks = [_BlockW(b) for b in ce.blocks]
                self.branches = []
            else:
                for bw in self.ce_blocks:
                    bw.refresh()
            for bi, est in enumerate(m.estimators):
                d = est.decoder
                C, nin = d.channels, d.in_channels
                ldp = _ceil(nin, 4)
                cc = d.cond_mlp[0].weight.shape[1]
                ch = d.cond_mlp[0].weight.shape[0]
                nl = len(d.blocks)
                te = d.time_mlp[0].weight.shape[1]
                if first:
                    br = type("BranchW", (), {})()
                    br.C, br.nin, br.ldp = C, nin, ldp
                    br.n_fft, br.hop = est.fft.n_fft, est.fft.hop_length
                    br.factor = est.cond_upsample_factor
                    br.dec = d
                    br.cc, br.ch, br.nl, br.te = cc, ch, nl, te
                    br.Win = br.Wout = br.cmW0 = br.cmW2 = None
                    dev = d.in_proj.weight.device
                    br.Wcp = torch.zeros(nl * C, cc, device=dev)
                    br.bcp = torch.zeros(nl * C, device=dev)
                    br.Wte = torch.zeros(nl * C, te, device=dev)
                    br.bte = torch.zeros(nl * C, device=dev)
                    br._halves = {}
                    self.branches.append(br)
                br = self.branches[bi]
                br.Win = _pack(d.in_proj.weight.detach(), C, nin, nin, 1, ldp, dst=br.Win)
                br.Wout = _pack(d.out_proj.weight.detach(), nin, C, C, 1, C, dst=br.Wout)
                br.cmW0 = _pack(d.cond_mlp[0].weight.detach(), ch, cc, cc, 1, cc, dst=br.cmW0)
                br.cmW2 = _pack(d.cond_mlp[2].weight.detach(), cc, ch, ch, 1, ch, dst=br.cmW2)
                for i, blk in enumerate(d.blocks):
                    _pack(blk.cond_proj.weight.detach(), C, cc, cc, 1, cc, dst=br.Wcp, dst_row=i * C)
                    _pack(blk.time_embed_proj.weight.detach(), C, te, te, 1, te, rnd=0, dst=br.Wte,
                          dst_row=i * C)
                    br.bcp[i * C:(i + 1) * C].copy_(blk.cond_proj.bias.detach())
                    br.bte[i * C:(i + 1) * C].copy_(blk.time_embed_proj.bias.detach())
                if first:
                    br.blocks = [_BlockW(b) for b in d.blocks]
                else:
                    for bw in br.blocks:
                        bw.refresh()
                for nm_, t in br._halves.items():       # fp16 copies already handed out: rewrite in place
                    t.copy_(getattr(br, nm_))
        self._built = True
        self.version += 1
        self.signature = _params_signature(self._plist)


def _g1(bw: _BlockW, a, h, M, bn=None, done=None, sat=None):
    bn = bn or BN_G1
    b = bw.blk
    if a.dtype == torch.float16:          # a1 (fp16) x W1 (fp16) -> h (fp16)
        return L.gemm_desc(a.data_ptr(), bw.half()[0].data_ptr(), h.data_ptr(), M, bw.H, bw.C, bw.C, bw.C,
                           bw.H, bn=bn, bias=b.pwconv1.bias.data_ptr(), slope=b.act.weight.data_ptr(),
                           act=L.ACT_PRELU, ab_f16=1, c_f16=1, done_counter=done, sat_flag=L.ptr(sat))
    return L.gemm_desc(a.data_ptr(), bw.W1.data_ptr(), h.data_ptr(), M, bw.H, bw.C, bw.C, bw.C, bw.H,
                       bn=bn, bias=b.pwconv1.bias.data_ptr(), slope=b.act.weight.data_ptr(),
                       act=L.ACT_PRELU, round_tf32=1)


# N tile of the chained pwconv2 problems by output width: the LPT tile schedule places these problems
# last, narrower tiles shorten the tail of the launch (measured: step 0.696 -> 0.672 ms, output equal to
# 1e-6, profiles/r02_switches.md)
_TAIL_BN = {768: 256, 512: 192, 384: 128}


def _g2(bw: _BlockW, h, x, M, bn=None, round_out=0, wait=None):
    bn = bn or (_TAIL_BN.get(bw.C) if wait is not None else None) or BN_G2
    b = bw.blk
    f16 = h.dtype == torch.float16        # h (fp16) x W2 (fp16) -> x (fp32 residual stream, in place)
    return L.gemm_desc(h.data_ptr(), (bw.half()[1] if f16 else bw.W2).data_ptr(), x.data_ptr(), M, bw.C,
                       bw.H, bw.H, bw.H, bw.C,
                       bn=bn, bias=b.pwconv2.bias.data_ptr(), res=x.data_ptr(), ld_res=bw.C,
                       res_scale=b.residual_scale.scale.data_ptr(), round_tf32=round_out, ab_f16=int(f16),
                       wait_counter=wait)


def _half(holder, name: str) -> Tensor:
    """fp16 copy of a packed (TF32-rounded, hence exactly representable) fp32 matrix of a branch;
    allocated once, kept current by PackedGenerator.refresh() (in place)."""
    t = holder._halves.get(name)
    if t is None:
        t = getattr(holder, name).to(torch.float16)
        holder._halves[name] = t
    return t


class InferencePlan:
    """Workspaces + launch sequence for one (B, mel frames, T, masked?) shape."""

    def __init__(self, model, packed: PackedGenerator, B: int, Fm: int, T: int, masked: bool):
        self.m, self.pk = model, packed
        self.B, self.Fm, self.T, self.masked = B, Fm, T, masked
        dev = next(model.parameters()).device
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        self.operands = getattr(model, "_block_operands", None) or BLOCK_OPERANDS
        op_dt = torch.float16 if self.operands == "f16" else torch.float32
        zo = lambda *s: torch.zeros(*s, device=dev, dtype=op_dt)      # block GEMM operands (a1, h)
        ce = model.cond_encoder
        self.n_mels = ce.cond_dim
        self.Cc = ce.channels
        self.Rc = B * Fm + 1                      # + one all-zero conditioning row
        self.mel = z(B, self.n_mels, Fm)
        self.mel_cl = z(B * Fm, packed.ld_mel)
        self.c0 = z(self.Rc, self.Cc)
        self.ce_a1 = zo(self.Rc, self.Cc)
        self.ce_h = zo(self.Rc, ce.blocks[0].hidden_channels)
        self.x_audio = z(B, T)
        self.lens = torch.zeros(B, device=dev, dtype=torch.int32)
        self.freqs = model.estimators[0].decoder.time_embed.freqs(dev)
        self.te_dim = model.estimators[0].decoder.time_embed.dim
        self.emb = z(1, self.te_dim)
        self.br = []
        for bw in packed.branches:
            w = type("BranchWS", (), {})()
            w.F = 1 + T // bw.hop
            w.R = B * w.F
            w.pin = z(w.R, bw.ldp)
            w.x = z(w.R, bw.C)
            w.a1 = zo(w.R, bw.C)
            w.h = zo(w.R, bw.blocks[0].H)
            w.pout = z(w.R, bw.ldp)
            w.fr = z(w.R, bw.n_fft)
            # the sampler evaluates every batch element at the same t (generator.py:260): the time
            # path is computed for ONE row and broadcast (row stride 0) to the block prologue
            w.te1 = z(1, bw.dec.time_mlp[0].weight.shape[0])
            w.te2 = z(1, bw.te)
            w.ts = z(1, bw.nl * bw.C)
            w.cm_h = zo(self.Rc, bw.ch)       # cond_mlp hidden / output: GEMM operands only
            w.c1 = zo(self.Rc, bw.cc)
            w.cp = z(self.Rc, bw.nl * bw.C)
            w.mask = z(w.R) if masked else None
            self.br.append(w)
        # chained pwconv1 -> pwconv2 launches: one counter per 256-row tile (cleared by block_pre)
        n_ce = (self.Rc + 255) // 256
        self._flags = torch.zeros(n_ce + 1, device=dev, dtype=torch.int32)       # [CondEncoder chain counters | range flag]
        self.ce_chain = self._flags[:n_ce]
        self.f16 = self.operands == "f16"
        self.chained = CHAIN_MLP and self.f16
        # fp16 range guard: every kernel that converts an operand to fp16 (block prologue: bit 1, GEMM
        # epilogues with an fp16 destination: bit 0) ORs into this flag when a value leaves +-65504 or is
        # not finite; it is cleared at the start of every launch sequence and mirrored to pinned host
        # memory behind it (generator.py reads it and falls back to TF32 operands)
        self.sat = self._flags[n_ce:]                 # cleared by the first CondEncoder prologue launch of a run
        self.sat_host = torch.zeros(1, dtype=torch.int32).pin_memory() if dev.type == "cuda" else torch.zeros(1, dtype=torch.int32)
        self.c0h = zo(self.Rc, self.Cc)
        self.cm_chain = torch.zeros(len(packed.branches) * ((self.Rc + 255) // 256), device=dev, dtype=torch.int32)
        offs, tot = [], 0
        for w in self.br:
            offs.append(tot)
            tot += (w.R + 255) // 256
        self.chain = torch.zeros(tot, device=dev, dtype=torch.int32)
        self.chain_off = offs
        self.t_all: Optional[Tensor] = None
        self._ts_cache: Dict[int, list] = {}        # F2G_CACHE_TIME: n -> per step -> per branch (1, nl*C)
        self._side: Optional[torch.cuda.Stream] = None
        self.graphs: Dict[Tuple[int, bool], torch.cuda.CUDAGraph] = {}
        self._seen: Dict[Tuple[int, bool], bool] = {}
        self._range_checked = False                 # generator.py: fp16 range flag read back after the first call

    # ------------------------------------------------------------------ step-invariant part
    def encode_cond(self) -> None:
        m, pk, B, Fm = self.m, self.pk, self.B, self.Fm
        ce = m.cond_encoder
        M = B * Fm
        L.im2col_cf(self.mel, B, self.n_mels, Fm, 3, self.mel_cl, pk.ld_mel, 1)
        L.gemm_group([L.gemm_desc(self.mel_cl.data_ptr(), pk.ce_Win.data_ptr(), self.c0.data_ptr(),
                                  M, self.Cc, 3 * self.n_mels, pk.ld_mel, pk.ld_mel, self.Cc,
                                  bias=ce.in_proj.bias.data_ptr())])
        L.biasnorm(self.c0, M, self.Cc, self.Cc, ce.in_norm.bias, ce.in_norm.log_scale, self.c0, self.Cc)
        for li, bw in enumerate(pk.ce_blocks):
            b = bw.blk
            last = li == len(pk.ce_blocks) - 1      # c0 is then only a GEMM operand: RN-round it
            pre = L.block_pre_desc(self.c0, B, Fm, bw.C, bw.C, bw.dwT, b.dwconv.bias, b.norm.bias,
                                   b.norm.log_scale, None, None, 0, 0, 1, 0, None, 0, self.ce_a1, bw.C,
                                   sat_flag=self.sat)
            if self.chained:
                # the first prologue of the run also clears the range flag behind the counters
                L.block_pre_group([pre], zero=self._flags if li == 0 else self.ce_chain)
                cnt = self.ce_chain.data_ptr()
                L.gemm_group([_g1(bw, self.ce_a1, self.ce_h, M, done=cnt, sat=self.sat),
                              _g2(bw, self.ce_h, self.c0, M, round_out=int(last), wait=cnt)])
            else:
                L.block_pre_group([pre])
                L.gemm_group([_g1(bw, self.ce_a1, self.ce_h, M, sat=self.sat)])
                L.gemm_group([_g2(bw, self.ce_h, self.c0, M, round_out=int(last))])
        self.cond_paths()

    def cond_paths(self) -> None:
        """cond_mlp + all cond_proj of every branch at mel-frame rate (+ the zero row)."""
        pk, Rc = self.pk, self.Rc
        if self.f16:
            # fp16 operands (c0 was RN-rounded to 11 significant bits by the last CondEncoder block, so
            # the conversion is exact); cond_mlp[0] -> PReLU -> cond_mlp[2] is one chained launch
            self.c0h.copy_(self.c0)                # torch's fp32 -> fp16 conversion overflows to inf: the first
            if self.chained:                       # consumer GEMM's epilogue then reports a non-finite value
                self.cm_chain.zero_()
            g0, g2, g3 = [], [], []
            ntile = (Rc + 255) // 256
            for i, (bw, w) in enumerate(zip(pk.branches, self.br)):
                cm = bw.dec.cond_mlp
                cnt = self.cm_chain.data_ptr() + 4 * i * ntile if self.chained else None
                W0, W2, Wc = _half(bw, "cmW0"), _half(bw, "cmW2"), _half(bw, "Wcp")
                g0.append(L.gemm_desc(self.c0h.data_ptr(), W0.data_ptr(), w.cm_h.data_ptr(), Rc, bw.ch, bw.cc,
                                      bw.cc, bw.cc, bw.ch, bias=cm[0].bias.data_ptr(),
                                      slope=cm[1].weight.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1,
                                      done_counter=cnt, sat_flag=L.ptr(self.sat)))
                # bias-only epilogue with an fp16 destination = "leaky ReLU with slope 1"
                g2.append(L.gemm_desc(w.cm_h.data_ptr(), W2.data_ptr(), w.c1.data_ptr(), Rc, bw.cc, bw.ch,
                                      bw.ch, bw.ch, bw.cc, bias=cm[2].bias.data_ptr(), act=L.ACT_LEAKY,
                                      leaky=1.0, ab_f16=1, c_f16=1, wait_counter=cnt, sat_flag=L.ptr(self.sat)))
                N = bw.nl * bw.C
                g3.append(L.gemm_desc(w.c1.data_ptr(), Wc.data_ptr(), w.cp.data_ptr(), Rc, N, bw.cc,
                                      bw.cc, bw.cc, N, bias=bw.bcp.data_ptr(), ab_f16=1))
            if self.chained:
                L.gemm_group(g0 + g2)
            else:
                L.gemm_group(g0)
                L.gemm_group(g2)
            L.gemm_group(g3)
            return
        ds = []
        for bw, w in zip(pk.branches, self.br):
            cm = bw.dec.cond_mlp
            ds.append(L.gemm_desc(self.c0.data_ptr(), bw.cmW0.data_ptr(), w.cm_h.data_ptr(), Rc, bw.ch,
                                  bw.cc, bw.cc, bw.cc, bw.ch, bias=cm[0].bias.data_ptr(),
                                  slope=cm[1].weight.data_ptr(), act=L.ACT_PRELU, round_tf32=1))
        L.gemm_group(ds)
        ds = []
        for bw, w in zip(pk.branches, self.br):
            cm = bw.dec.cond_mlp
            ds.append(L.gemm_desc(w.cm_h.data_ptr(), bw.cmW2.data_ptr(), w.c1.data_ptr(), Rc, bw.cc,
                                  bw.ch, bw.ch, bw.ch, bw.cc, bias=cm[2].bias.data_ptr(), round_tf32=1))
        L.gemm_group(ds)
        ds = []
        for bw, w in zip(pk.branches, self.br):
            N = bw.nl * bw.C
            ds.append(L.gemm_desc(w.c1.data_ptr(), bw.Wcp.data_ptr(), w.cp.data_ptr(), Rc, N, bw.cc,
                                  bw.cc, bw.cc, N, bias=bw.bcp.data_ptr()))
        L.gemm_group(ds)

    # ------------------------------------------------------------------ one model evaluation
    def process_model(self, t_dev: Tensor) -> None:
        """x_audio, cond (cp) -> per-branch windowed iSTFT frames (w.fr).  t_dev: (B,) device."""
        self.process_front(t_dev)
        self.process_blocks()

    def process_front(self, t_dev: Tensor, with_time: bool = True) -> None:
        """The part of one model evaluation that does not read the conditioning: STFT, in_proj,
        in_norm and (unless the step's vectors are cached) the time-embedding path."""
        pk, B, T, Fm = self.pk, self.B, self.T, self.Fm
        L.stft_group([(self.x_audio, w.pin, bw.n_fft, bw.hop, w.F, w.R, T, bw.ldp)
                      for bw, w in zip(pk.branches, self.br)], B, T, round_tf32=1)
        L.gemm_group([L.gemm_desc(w.pin.data_ptr(), bw.Win.data_ptr(), w.x.data_ptr(), w.R, bw.C,
                                  bw.nin, bw.ldp, bw.ldp, bw.C, bias=bw.dec.in_proj.bias.data_ptr())
                      for bw, w in zip(pk.branches, self.br)])
        for bw, w in zip(pk.branches, self.br):
            L.biasnorm(w.x, w.R, bw.C, bw.C, bw.dec.in_norm.bias, bw.dec.in_norm.log_scale, w.x, bw.C)
        if with_time:
            self.time_path(t_dev)

    def time_path(self, t_dev: Tensor) -> None:
        """SinusoidalPosEmb -> time_mlp -> the 8 time_embed_proj of every branch, for ONE row (all batch
        elements share t at inference); writes w.ts."""
        pk = self.pk
        L.time_sinusoid(t_dev, 1, self.te_dim, self.freqs, 1000.0, self.emb)
        prob1, prob2, prob3 = [], [], []
        for bw, w in zip(pk.branches, self.br):
            tm = bw.dec.time_mlp
            H = tm[0].weight.shape[0]
            prob1.append((self.emb, bw.te, self.te_dim, tm[0].weight, bw.te, tm[0].bias, H, w.te1, H))
            prob2.append((w.te1, H, H, tm[2].weight, H, tm[2].bias, bw.te, w.te2, bw.te))
            prob3.append((w.te2, bw.te, bw.te, bw.Wte, bw.te, bw.bte, bw.nl * bw.C, w.ts, bw.nl * bw.C))
        L.linear_small_group(prob1, 1, L.ACT_SILU)
        L.linear_small_group(prob2, 1, L.ACT_NONE)
        L.linear_small_group(prob3, 1, L.ACT_NONE)

    def process_blocks(self) -> None:
        """ConvNeXt blocks (read cp, the per-layer conditioning rows), out_proj, inverse FFT."""
        pk, B, T, Fm = self.pk, self.B, self.T, self.Fm
        nl = pk.branches[0].nl
        for i in range(nl):
            pre = []
            for bw, w in zip(pk.branches, self.br):
                k = bw.blocks[i]
                b = k.blk
                ldc = bw.nl * bw.C
                pre.append(L.block_pre_desc(w.x, B, w.F, bw.C, bw.C, k.dwT, b.dwconv.bias, b.norm.bias,
                                            b.norm.log_scale, w.mask, w.cp[:, i * bw.C:], ldc, Fm,
                                            bw.factor, B * Fm, w.ts[:, i * bw.C:], 0, w.a1, bw.C,
                                            sat_flag=self.sat))
            if self.chained:                # prologues, then pwconv1 -> pwconv2 of all branches: 2 launches
                L.block_pre_group(pre, zero=self.chain)
                cnt = [self.chain.data_ptr() + 4 * o for o in self.chain_off]
                L.gemm_group([_g1(bw.blocks[i], w.a1, w.h, w.R, done=c, sat=self.sat)
                              for bw, w, c in zip(pk.branches, self.br, cnt)] +
                             [_g2(bw.blocks[i], w.h, w.x, w.R, round_out=int(i == nl - 1), wait=c)
                              for bw, w, c in zip(pk.branches, self.br, cnt)])
                continue
            L.block_pre_group(pre)          # the three branches' prologues: one launch
            L.gemm_group([_g1(bw.blocks[i], w.a1, w.h, w.R, sat=self.sat) for bw, w in zip(pk.branches, self.br)])
            L.gemm_group([_g2(bw.blocks[i], w.h, w.x, w.R, round_out=int(i == nl - 1))
                          for bw, w in zip(pk.branches, self.br)])
        L.gemm_group([L.gemm_desc(w.x.data_ptr(), bw.Wout.data_ptr(), w.pout.data_ptr(), w.R, bw.nin,
                                  bw.C, bw.C, bw.C, bw.ldp, bias=bw.dec.out_proj.bias.data_ptr(),
                                  row_scale=L.ptr(w.mask))
                      for bw, w in zip(pk.branches, self.br)])
        L.irfft_group([(w.pout, w.fr, bw.n_fft, 0, 0, w.R, bw.ldp, bw.n_fft)
                       for bw, w in zip(pk.branches, self.br)])

    def combine(self, out: Tensor, euler: bool, t: float, dt: float, clamp: bool,
                weight: Optional[Tensor] = None) -> None:
        pk = self.pk
        L.ola_combine([w.fr for w in self.br], [bw.n_fft for bw in pk.branches],
                      [bw.hop for bw in pk.branches], [w.F for w in self.br], weight, self.x_audio,
                      out, self.B, self.T, euler, t, dt, clamp)

    # ------------------------------------------------------------------ Euler sampler
    def _prepare_steps(self, n: int):
        t_span = torch.linspace(0, 1, n + 1)                       # generator.py:253
        dt = float(t_span[1] - t_span[0])
        ts = [float(t_span[k]) for k in range(n)]
        # built on the device (no host->device copy: this also runs under stream capture)
        self.t_all = torch.linspace(0, 1, n + 1, device=self.x_audio.device)[:n].unsqueeze(1) \
            .expand(n, self.B).contiguous()
        self._refresh_time_cache(n)
        return ts, dt

    def _refresh_time_cache(self, n: int) -> None:
        """Per-step time-scale vectors of an n-step sampler, recomputed IN PLACE whenever the packed
        weights changed (PackedGenerator.version) -- captured graphs keep pointing at the same
        tensors.  Under an outer stream capture (the trainer's phase graphs) they are always
        recomputed, so that the outer graph carries its own time path."""
        if not CACHE_TIME:
            return
        ent = self._ts_cache.get(n)
        if ent is not None and ent[0] == self.pk.version and not torch.cuda.is_current_stream_capturing():
            return
        per_step = ent[1] if ent is not None else [[torch.zeros_like(w.ts) for w in self.br] for _ in range(n)]
        t_rows = torch.linspace(0, 1, n + 1, device=self.x_audio.device)[:n].contiguous()
        keep = [w.ts for w in self.br]
        for k in range(n):
            for w, t in zip(self.br, per_step[k]):
                w.ts = t
            self.time_path(t_rows[k:k + 1])
        for w, t in zip(self.br, keep):
            w.ts = t
        self._ts_cache[n] = (self.pk.version, per_step)

    def _run(self, n: int, clamp: bool, with_cond: bool = True) -> None:
        ts, dt = self._steps
        if self.f16 and not (with_cond and self.chained):
            self.sat.zero_()                         # otherwise cleared by the CondEncoder's first prologue launch
        forked = with_cond and FORK_COND
        if forked:
            # The conditioning path (CondEncoder + cond_mlp + cond_proj: ~20 launches of GEMMs with
            # 36-72 tiles, i.e. at most half the SMs busy) and the front of the first model
            # evaluation (STFT, in_proj, norms, time MLP: small kernels) are independent: run them
            # on two streams and join before the first block prologue.  Under stream capture the
            # fork/join becomes two parallel branches of the graph.
            if self.f16:                       # lazily built fp16 weights: allocate on the main stream
                for bw in self.pk.ce_blocks:
                    bw.half()
                for bw in self.pk.branches:
                    for nm in ("cmW0", "cmW2", "Wcp"):
                        _half(bw, nm)
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.x_audio.device)
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            self._side.wait_event(fork)
            with torch.cuda.stream(self._side):
                self.encode_cond()
                join.record(self._side)
            self.process_front(self.t_all[0], with_time=not CACHE_TIME)
            main.wait_event(join)
        elif with_cond:
            self.encode_cond()
        cached = self._ts_cache[n][1] if CACHE_TIME else None
        for k in range(n):
            if not (forked and k == 0):
                self.process_front(self.t_all[k], with_time=cached is None)
            if cached is not None:
                for w, ts_k in zip(self.br, cached[k]):
                    w.ts = ts_k
            self.process_blocks()
            self.combine(self.x_audio, True, ts[k], dt, clamp and k == n - 1)

    def set_masks(self, lens: Tensor) -> None:
        self.lens.copy_(lens.to(torch.int32))
        for bw, w in zip(self.pk.branches, self.br):
            L.frame_mask(self.lens, self.B, w.F, bw.hop, w.mask)

    def infer(self, mel: Tensor, noise: Optional[Tensor], lens: Optional[Tensor], n: int, clamp: bool,
              use_graph: bool = True, noise_scale: float = 1.0, out: Optional[Tensor] = None) -> Tensor:
        """`mel` may live on the host (pinned or not): it is copied straight into the plan's static
        input buffer.  `noise=None` draws randn * noise_scale from torch's global CUDA generator
        directly into the sample buffer (same draw as the reference's `torch.randn(...) * scale`,
        generator.py:356).  `out` (optional, host or device) receives the audio instead of a fresh
        device tensor."""
        self.mel.copy_(mel, non_blocking=True)
        if noise is None:
            # one kernel; torch.randn IS empty().normal_(0, 1) and normal_(0, s) applies `z * s + 0`
            # to the same Philox stream, so this equals the reference's `torch.randn(...) * s`
            self.x_audio.normal_(0.0, noise_scale)
        else:
            self.x_audio.copy_(noise, non_blocking=True)
        if self.masked:
            self.set_masks(lens)
        key = (n, bool(clamp))

        def result() -> Tensor:
            if self.f16 and self.sat_host.is_pinned() and not torch.cuda.is_current_stream_capturing():
                # the range flag follows every launch sequence to pinned host memory (4 bytes, outside the
                # graph: a D2H node would sit on the replay's critical path); generator.py reads it
                self.sat_host.copy_(self.sat, non_blocking=True)
            if out is None:
                return self.x_audio.clone()
            out.copy_(self.x_audio, non_blocking=True)
            return out

        dev = self.x_audio.device
        if not use_graph:
            self._steps = self._prepare_steps(n)
            with _chain_guard(dev):
                self._run(n, clamp)
            return result()
        g = self.graphs.get(key)
        if g is None and not self._seen.get(key):
            # first call for this shape: run eagerly; capture on the second call
            self._seen[key] = True
            self._steps = self._prepare_steps(n)
            with _chain_guard(dev):
                self._run(n, clamp)
            return result()
        if g is None:
            x0 = self.x_audio.clone()
            self._steps = self._prepare_steps(n)
            with _chain_guard(dev):
                self._run(n, clamp)                 # warm-up (also sets kernel attributes)
            self.x_audio.copy_(x0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(n, clamp)
            self.graphs[key] = (g, self._steps, self.t_all)
        else:
            g, self._steps, self.t_all = g
            self._refresh_time_cache(n)             # weights repacked in place since the capture?
        with _chain_guard(dev):
            g.replay()
        return result()

    def infer_from_cond(self, cond: Tensor, noise: Tensor, lens: Optional[Tensor], n: int,
                        clamp: bool) -> Tensor:
        """BaseAudioGenerator.infer entry: `cond` is an already-encoded (B, C, Fm) tensor."""
        self.c0[: self.B * self.Fm].copy_(cond.transpose(1, 2).reshape(self.B * self.Fm, self.Cc))
        self.x_audio.copy_(noise)
        if self.masked:
            self.set_masks(lens)
        self._steps = self._prepare_steps(n)
        with _chain_guard(self.x_audio.device):
            self.cond_paths()
            self._run(n, clamp, with_cond=False)
        return self.x_audio.clone()
