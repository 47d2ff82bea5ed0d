"""Loss terms of the hot path with CUDA forward/backward: filterbank spectrograms (STFT + |.|^p +
triangular filters [+ safe_log]) as one fused kernel each way, and the reductions built on them:
  * spectral-energy-scaled flow-matching loss   (flow2gan/models/generator.py:172-200)
  * multi-scale log-mel L1 reconstruction loss  (flow2gan/models/gan.py:89-99)
The scalar reductions over the (small) filterbank tensors are plain element-wise torch ops."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib as L


class _FilterSpecFn(torch.autograd.Function):
    """(B, T) -> (B*frames, n_filt) rows: [log] fb^T |STFT|^p, differentiable w.r.t. the audio."""

    @staticmethod
    def forward(ctx, audio: Tensor, fb: Tensor, n_fft: int, hop: int, mode: int, log_clip: float):
        audio = audio.contiguous().float()
        B, T = audio.shape
        frames = 1 + T // hop
        n_filt = fb.shape[1]
        out = torch.empty(B * frames, n_filt, device=audio.device, dtype=torch.float32)
        L.stft(audio, B, T, T, n_fft, hop, mode, out, n_filt, fb=fb, n_filt=n_filt, log_clip=log_clip)
        ctx.save_for_backward(audio, fb)
        ctx.cfg = (n_fft, hop, mode, log_clip, B, T, frames, n_filt)
        return out

    @staticmethod
    def backward(ctx, dF: Tensor):
        audio, fb = ctx.saved_tensors
        n_fft, hop, mode, log_clip, B, T, frames, n_filt = ctx.cfg
        dF = dF.contiguous()
        fr = torch.empty(B * frames, n_fft, device=audio.device, dtype=torch.float32)
        L.spec_loss_bwd(audio, B, T, T, n_fft, hop, mode, fb, n_filt, log_clip, dF, n_filt, fr)
        dx = torch.empty(B, T, device=audio.device, dtype=torch.float32)
        L.stft_bwd_fold(fr, B, T, n_fft, hop, frames, dx, False)
        return dx, None, None, None, None, None


def filter_spec_rows(audio: Tensor, fb: Tensor, n_fft: int, hop: int, mode: int, log_clip: float = 0.0) -> Tensor:
    return _FilterSpecFn.apply(audio, fb.contiguous(), n_fft, hop, mode, float(log_clip))


def spectral_scaled_loss(model, pred: Tensor, ref: Tensor, audio_lens: Tensor) -> Tensor:
    """BaseAudioGenerator.compute_loss, spec_scaling_loss branch (generator.py:179-200)."""
    ls = model.loss_spec
    B, T = ref.shape
    frames = 1 + T // ls.hop_length
    err = pred - ref
    with torch.no_grad():
        gt = filter_spec_rows(ref, ls.fb, ls.n_fft, ls.hop_length, L.SPEC_POWER).view(B, frames, -1)
        scale = ((gt + model.loss_eps) ** -model.loss_power).clamp(min=model.loss_scale_min,
                                                                  max=model.loss_scale_max)
        spec_lens = torch.div(audio_lens, ls.hop_length, rounding_mode="floor") + 1
        mask = (torch.arange(frames, device=ref.device)[None, :] < spec_lens[:, None]).unsqueeze(-1)
    es = filter_spec_rows(err, ls.fb, ls.n_fft, ls.hop_length, L.SPEC_POWER).view(B, frames, -1)
    return (es * scale * mask).sum() / (mask.sum() * es.shape[2])


def mel_recon_loss(mel_modules, real: Tensor, fake: Tensor, fused: bool = False) -> Tensor:
    """GAN.mel_recon_loss (gan.py:89-99): sum_k L1(log-mel_k(real), log-mel_k(fake))."""
    loss = 0
    rs, fs = [], []
    for m in mel_modules:
        fb = m.mel_scale.fb
        with torch.no_grad():
            r = filter_spec_rows(real, fb, m.n_fft, m.hop_length, L.SPEC_MAG, 1e-7)
        f = filter_spec_rows(fake, fb, m.n_fft, m.hop_length, L.SPEC_MAG, 1e-7)
        if fused:
            rs.append(r)
            fs.append(f)
        else:
            loss = loss + torch.nn.functional.l1_loss(r, f)
    return l1_terms(rs, fs) if fused else loss


# ---------------------------------------------------------------------------------------------
# fused multi-tensor reductions of GAN.forward's loss terms (csrc/losses.cu): one launch per <= 24
# terms each way instead of 3-4 element-wise torch launches per term and direction
# ---------------------------------------------------------------------------------------------
class _L1TermsFn(torch.autograd.Function):
    """sum_i mean|ref_i - x_i|, differentiable w.r.t. the x_i (refs are constants)."""

    @staticmethod
    def forward(ctx, n: int, *tensors: Tensor):
        refs, xs = tensors[:n], tensors[n:]
        out = torch.empty(1, device=xs[0].device, dtype=torch.float32)
        L.loss_terms([L.loss_term(L.LOSS_L1, r, x, None) for r, x in zip(refs, xs)], False, out, None)
        ctx.save_for_backward(*tensors)
        ctx.n = n
        return out[0]

    @staticmethod
    def backward(ctx, g: Tensor):
        n = ctx.n
        refs, xs = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        grads = [torch.empty(x.shape, device=x.device, dtype=torch.float32) for x in xs]
        gout = g.reshape(1).float().contiguous()
        L.loss_terms([L.loss_term(L.LOSS_L1, r, x, d) for r, x, d in zip(refs, xs, grads)], True, None, gout)
        return (None,) + (None,) * n + tuple(grads)


class _HingeTermsFn(torch.autograd.Function):
    """sum_i mean(clamp(1 + sign_i * s_i, min=0)), differentiable w.r.t. the scores."""

    @staticmethod
    def forward(ctx, signs, *scores: Tensor):
        out = torch.empty(1, device=scores[0].device, dtype=torch.float32)
        L.loss_terms([L.loss_term(L.LOSS_HINGE, s, None, None, sg) for s, sg in zip(scores, signs)], False, out, None)
        ctx.save_for_backward(*scores)
        ctx.signs = tuple(signs)
        return out[0]

    @staticmethod
    def backward(ctx, g: Tensor):
        scores = ctx.saved_tensors
        grads = [torch.empty(s.shape, device=s.device, dtype=torch.float32) for s in scores]
        gout = g.reshape(1).float().contiguous()
        L.loss_terms([L.loss_term(L.LOSS_HINGE, s, None, d, sg) for s, d, sg in zip(scores, grads, ctx.signs)],
                     True, None, gout)
        return (None,) + tuple(grads)


def _as4d(t: Tensor) -> Tensor:
    return t if t.dim() <= 4 else t.reshape(-1, *t.shape[-3:])


def l1_terms(refs, xs) -> Tensor:
    """sum_i l1_loss(refs[i].detach(), xs[i]) (feature_matching_loss gan.py:72-87, mel_recon_loss :89-99)."""
    refs = [_as4d(r.detach().float()) for r in refs]
    xs = [_as4d(x.float()) for x in xs]
    return _L1TermsFn.apply(len(refs), *refs, *xs)


def hinge_terms(scores, signs) -> Tensor:
    """sum_i mean(clamp(1 + signs[i] * scores[i], min=0)) (discriminator_loss gan.py:57-63 with signs
    -1 for real / +1 for fake scores; generator_loss :65-70 with -1 for fake scores)."""
    return _HingeTermsFn.apply(tuple(float(s) for s in signs), *[_as4d(s.float()) for s in scores])
