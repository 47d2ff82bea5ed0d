"""Parameter containers with the reference's module tree (=> identical state_dict keys and
shapes) plus the front-end modules that run entirely on the CUDA library.

The nn.Conv1d / nn.Linear / nn.PReLU instances below are used as *parameter holders only*:
their torch forward is never called on the hot path -- the math runs in csrc/*.cu through
flow2gan_b200.engine (inference) / flow2gan_b200.train (fwd+bwd).
Mirrors flow2gan/models/modules.py (class names, constructor arguments, buffers).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib as L


def melscale_fbanks(n_freqs: int, n_mels: int, sample_rate: int) -> Tensor:
    """HTK mel triangular filters, norm=None, f_min=0, f_max=sr//2 (the torchaudio
    MelSpectrogram defaults the reference relies on, modules.py:131-138) -> (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + float(sample_rate // 2) / 700.0)
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    return _triangles(all_freqs, f_pts)


def linear_fbanks(n_freqs: int, n_filter: int, sample_rate: int) -> Tensor:
    """Linear-frequency triangular filters (modules.py:194-200) -> (n_freqs, n_filter)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    f_pts = torch.linspace(0.0, float(sample_rate // 2), n_filter + 2)
    return _triangles(all_freqs, f_pts)


def _triangles(all_freqs: Tensor, f_pts: Tensor) -> Tensor:
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


class _Window(nn.Module):
    """Holder for an STFT window buffer (state_dict keys `*.window`)."""

    def __init__(self, n_fft: int, hop_length: int):
        super().__init__()
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = n_fft
        self.register_buffer("window", torch.hann_window(n_fft))


class STFT(_Window):
    pass


class ISTFT(_Window):
    pass


class _MelScale(nn.Module):
    def __init__(self, fb: Tensor):
        super().__init__()
        self.register_buffer("fb", fb)


def filtered_spectrogram(waveform: Tensor, n_fft: int, hop: int, mode: int, fb: Tensor,
                         log_clip: float) -> Tensor:
    """(..., T) -> (..., n_filt, 1 + T//hop) through one fused f2g_stft launch (no autograd:
    front-end / target-side use)."""
    lead = waveform.shape[:-1]
    T = waveform.shape[-1]
    x = waveform.detach().reshape(-1, T).contiguous().float()
    B = x.shape[0]
    frames = 1 + T // hop
    n_filt = fb.shape[1]
    out = torch.empty(B * frames, n_filt, device=x.device, dtype=torch.float32)
    L.stft(x, B, T, T, n_fft, hop, mode, out, n_filt, fb=fb.contiguous(), n_filt=n_filt,
           log_clip=log_clip)
    return out.view(B, frames, n_filt).transpose(1, 2).reshape(*lead, n_filt, frames)


class MelSpectrogram(nn.Module):
    """torchaudio.transforms.MelSpectrogram(power=1, center=True) equivalent running on
    f2g_stft (state_dict keys: spectrogram.window, mel_scale.fb)."""

    def __init__(self, sample_rate: int, n_fft: int, hop_length: int, n_mels: int,
                 center: bool = True, power: float = 1):
        super().__init__()
        assert center and power == 1, "only the configuration the reference uses is built"
        self.n_fft, self.hop_length, self.n_mels = n_fft, hop_length, n_mels
        self.spectrogram = _Window(n_fft, hop_length)
        self.mel_scale = _MelScale(melscale_fbanks(n_fft // 2 + 1, n_mels, sample_rate))

    def forward(self, waveform: Tensor, log_clip: float = 0.0) -> Tensor:
        return filtered_spectrogram(waveform, self.n_fft, self.hop_length, L.SPEC_MAG,
                                    self.mel_scale.fb, log_clip)


class LogMelSpectrogram(nn.Module):
    """flow2gan/models/modules.py:119-143: log(clip(mel(|STFT|), 1e-7)), one fused kernel."""

    def __init__(self, sampling_rate: int = 24000, n_fft: int = 1024, hop_length: int = 256,
                 n_mels: int = 100, center: bool = True, power: float = 1):
        super().__init__()
        self.mel = MelSpectrogram(sampling_rate, n_fft, hop_length, n_mels, center, power)

    def forward(self, waveform: Tensor) -> Tensor:
        return self.mel(waveform, log_clip=1e-7)


class LinearFilterSpectrogram(nn.Module):
    """flow2gan/models/modules.py:146-214 (power-2 STFT -> linear triangular filters)."""

    def __init__(self, sample_rate: int, n_filter: int, n_fft: int,
                 hop_length: Optional[int] = None, center: bool = True, power: float = 2.0):
        super().__init__()
        assert center and power == 2.0
        self.sample_rate, self.n_fft, self.n_filter = sample_rate, n_fft, n_filter
        self.hop_length = hop_length if hop_length is not None else n_fft // 2
        self.spectrogram = _Window(n_fft, self.hop_length)
        self.register_buffer("fb", linear_fbanks(n_fft // 2 + 1, n_filter, sample_rate))

    def forward(self, waveform: Tensor) -> Tensor:
        return filtered_spectrogram(waveform, self.n_fft, self.hop_length, L.SPEC_POWER,
                                    self.fb, 0.0)


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0, "SinusoidalPosEmb requires dim to be even"
        self.dim = dim

    def freqs(self, device) -> Tensor:
        half = self.dim // 2
        e = math.log(10000) / (half - 1)
        return torch.exp(torch.arange(half, device=device).float() * -e)   # modules.py:227-229


class ChannelScale(nn.Module):
    def __init__(self, channels: int, scale: float = 1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.full((channels, 1), scale))


class BiasNorm(nn.Module):
    def __init__(self, num_channels: int, channel_dim: int = -1, log_scale: float = 1.0,
                 log_scale_min: float = -1.5, log_scale_max: float = 1.5):
        super().__init__()
        self.num_channels = num_channels
        self.channel_dim = channel_dim
        self.log_scale = nn.Parameter(torch.tensor(log_scale))
        self.bias = nn.Parameter(torch.empty(num_channels).normal_(mean=0, std=1e-2))
        self.log_scale_min = log_scale_min
        self.log_scale_max = log_scale_max


class ConvNeXtBlock(nn.Module):
    def __init__(self, channels: int = 512, hidden_channels: int = 1536,
                 conv_kernel_size: int = 7, cond_channels: Optional[int] = None,
                 time_embed_channels: Optional[int] = None,
                 residual_scale: Optional[float] = 1.0):
        super().__init__()
        if conv_kernel_size != 7:
            raise NotImplementedError("the fused block prologue is built for kernel size 7")
        if residual_scale is None:
            raise NotImplementedError("residual_scale=None is not used by any released config")
        self.channels, self.hidden_channels = channels, hidden_channels
        self.dwconv = nn.Conv1d(channels, channels, kernel_size=7, padding=3, groups=channels)
        self.norm = BiasNorm(channels, channel_dim=1)
        self.pwconv1 = nn.Conv1d(channels, hidden_channels, kernel_size=1)
        self.act = nn.PReLU(hidden_channels)
        self.pwconv2 = nn.Conv1d(hidden_channels, channels, kernel_size=1)
        if cond_channels is not None:
            self.cond_proj = nn.Conv1d(cond_channels, channels, kernel_size=1)
        if time_embed_channels is not None:
            self.time_embed_proj = nn.Linear(time_embed_channels, channels)
        self.residual_scale = ChannelScale(channels)


class CondEncoder(nn.Module):
    def __init__(self, cond_dim: int = 100, channels: int = 512, hidden_factor: int = 3,
                 conv_kernel_size: int = 7, num_layers: int = 4,
                 residual_scale: Optional[float] = 1.0):
        super().__init__()
        self.cond_dim, self.channels = cond_dim, channels
        self.in_proj = nn.Conv1d(cond_dim, channels, kernel_size=3, padding=1)
        self.in_norm = BiasNorm(channels, channel_dim=1)
        self.blocks = nn.ModuleList([
            ConvNeXtBlock(channels, int(channels * hidden_factor), conv_kernel_size,
                          residual_scale=residual_scale) for _ in range(num_layers)])


class ConvNeXtDecoder(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, channels: int = 512,
                 cond_channels: int = 512, time_embed_channels: int = 512,
                 hidden_factor: int = 3, conv_kernel_size: int = 7, num_layers: int = 8,
                 residual_scale: Optional[float] = 1.0, use_t: bool = True):
        super().__init__()
        if not use_t:
            raise NotImplementedError("use_t=False is not used by any released config")
        self.in_channels, self.out_channels, self.channels = in_channels, out_channels, channels
        self.in_proj = nn.Conv1d(in_channels, channels, kernel_size=1)
        self.in_norm = BiasNorm(channels, channel_dim=1)
        self.time_embed = SinusoidalPosEmb(time_embed_channels)
        hidden = int(time_embed_channels * hidden_factor)
        self.time_mlp = nn.Sequential(nn.Linear(time_embed_channels, hidden), nn.SiLU(),
                                      nn.Linear(hidden, time_embed_channels))
        cond_hidden = int(cond_channels * hidden_factor)
        self.cond_mlp = nn.Sequential(nn.Conv1d(cond_channels, cond_hidden, kernel_size=1),
                                      nn.PReLU(cond_hidden),
                                      nn.Conv1d(cond_hidden, cond_channels, kernel_size=1))
        self.blocks = nn.ModuleList([
            ConvNeXtBlock(channels, int(channels * hidden_factor), conv_kernel_size,
                          cond_channels, time_embed_channels, residual_scale)
            for _ in range(num_layers)])
        self.out_proj = nn.Conv1d(channels, out_channels, kernel_size=1)


class AudioConvNeXt(nn.Module):
    def __init__(self, n_fft: int = 512, hop_length: int = 256, cond_hop_length: int = 256,
                 channels: int = 768, cond_channels: int = 512, time_embed_channels: int = 512,
                 hidden_factor: int = 3, conv_kernel_size: int = 7, num_layers: int = 8,
                 residual_scale: Optional[float] = 1.0, use_t: bool = True):
        super().__init__()
        self.fft = STFT(n_fft=n_fft, hop_length=hop_length)
        self.ifft = ISTFT(n_fft=n_fft, hop_length=hop_length)
        assert cond_hop_length % hop_length == 0, \
            "cond_hop_length should be integer multiple of hop_length."
        self.cond_upsample_factor = cond_hop_length // hop_length
        self.decoder = ConvNeXtDecoder(n_fft + 2, n_fft + 2, channels, cond_channels,
                                       time_embed_channels, hidden_factor, conv_kernel_size,
                                       num_layers, residual_scale, use_t)
