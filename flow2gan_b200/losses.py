"""Loss terms of the hot path with CUDA forward/backward: filterbank spectrograms (STFT + |.|^p +
triangular filters [+ safe_log]) as one fused kernel each way, and the reductions built on them:
  * spectral-energy-scaled flow-matching loss   (flow2gan/models/generator.py:172-200)
  * multi-scale log-mel L1 reconstruction loss  (flow2gan/models/gan.py:89-99)
The scalar reductions over the (small) filterbank tensors are plain element-wise torch ops."""
from __future__ import annotations

from typing import Optional

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


def mel_recon_loss(mel_modules, real: Tensor, fake: Tensor) -> Tensor:
    """GAN.mel_recon_loss (gan.py:89-99): sum_k L1(log-mel_k(real), log-mel_k(fake))."""
    loss = 0
    for m in mel_modules:
        fb = m.mel_scale.fb
        with torch.no_grad():
            r = filter_spec_rows(real, fb, m.n_fft, m.hop_length, L.SPEC_MAG, 1e-7)
        f = filter_spec_rows(fake, fb, m.n_fft, m.hop_length, L.SPEC_MAG, 1e-7)
        loss = loss + torch.nn.functional.l1_loss(r, f)
    return loss
