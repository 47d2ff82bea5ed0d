"""CPU oracle for the Flow2GAN hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain, functional (no nn.Module) fp32 CPU restatement of the reference's
algorithm for the path named in BASELINE.json.  It exists so that the CUDA product path in
``flow2gan_b200`` can be checked against something that does not share any code with it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it.  The product package never does.

Pinning status (see DESIGN.md "Oracle"):
  * ``log_mel``  -- pinned by the reference's own wav<->mel fixtures (tests/golden/mel_*.pt).
  * everything else -- the reference has no tests; pinned against outputs of the *reference
    itself* imported in the build container (tests/golden/make_golden.py writes
    tests/golden/ref_*.pt; tests/test_oracle_vs_golden.py checks this file against them).

All functions take a flat ``sd`` (the reference's ``state_dict`` layout)
plus a key prefix.  Citations are into /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------
# configs (flow2gan/models/config.py:31-115)
# ----------------------------------------------------------------------------------------
def generator_config(name: str) -> dict:
    base = dict(
        channels=(768, 512, 384), time_embed_channels=512, hidden_factor=3,
        conv_kernel_sizes=(7, 7, 7), num_layers=(8, 8, 8), cond_enc_channels=512,
        cond_enc_hidden_factor=3, cond_enc_conv_kernel_size=7, cond_enc_num_layers=4,
        init_noise_scale=0.1, loss_n_filters=256, loss_power=0.5, loss_eps=1e-7,
        loss_scale_min=1e-2, loss_scale_max=1e2, branch_dropout=0.05,
    )
    if name == "mel_24k_base":
        base.update(sampling_rate=24000, n_mels=100, mel_n_fft=1024, mel_hop_length=256,
                    n_ffts=(512, 256, 128), hop_lengths=(256, 128, 64),
                    loss_n_fft=1024, loss_hop_length=256)
    elif name == "mel_44k_128band_512x_base":
        base.update(sampling_rate=44100, n_mels=128, mel_n_fft=2048, mel_hop_length=512,
                    n_ffts=(1024, 512, 256), hop_lengths=(512, 256, 128),
                    loss_n_fft=2048, loss_hop_length=512)
    else:
        raise ValueError(f"Unsupported model name: {name}")
    return base


# ----------------------------------------------------------------------------------------
# STFT / iSTFT (modules.py:52-116, torch.stft/istft semantics)
# ----------------------------------------------------------------------------------------
def hann(n: int) -> Tensor:
    # torch.hann_window(n) (periodic): 0.5 - 0.5 cos(2 pi k / n)
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).float()


def frame_signal(x: Tensor, n_fft: int, hop: int) -> Tensor:
    """center=True, reflect pad n/2 each side; frames f = 0..T//hop.  -> (B, F, n_fft)"""
    xp = F.pad(x.unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect").squeeze(1)
    return xp.unfold(-1, n_fft, hop)


def stft_complex(x: Tensor, n_fft: int, hop: int) -> Tensor:
    """modules.py:69-78.  x (B,T) -> complex (B, n/2+1, 1+T//hop), unnormalised, onesided."""
    fr = frame_signal(x, n_fft, hop) * hann(n_fft)
    return torch.fft.rfft(fr, dim=-1).transpose(1, 2)


def stft_packed(x: Tensor, n_fft: int, hop: int) -> Tensor:
    """STFT + fft_to_real (modules.py:31-38): (B, n+2, F) = [Re bins ; Im bins] planar."""
    s = stft_complex(x, n_fft, hop)
    return torch.cat([s.real, s.imag], dim=1)


def istft_packed(p: Tensor, n_fft: int, hop: int) -> Tensor:
    """real_to_fft + torch.istft(center=True) (modules.py:41-49,105-116).
    p (B, n+2, F) -> (B, hop*(F-1))."""
    B, _, Fr = p.shape
    nb = n_fft // 2 + 1
    spec = torch.complex(p[:, :nb], p[:, nb:])                  # (B, nb, F)
    w = hann(n_fft)
    fr = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * w   # (B, F, n)
    total = n_fft + hop * (Fr - 1)
    y = torch.zeros(B, total)
    env = torch.zeros(total)
    w2 = w * w
    for f in range(Fr):
        y[:, f * hop:f * hop + n_fft] += fr[:, f]
        env[f * hop:f * hop + n_fft] += w2
    s, e = n_fft // 2, n_fft // 2 + hop * (Fr - 1)
    return y[:, s:e] / env[s:e]


def convert_length(x: Tensor, length: int) -> Tensor:
    """utils.py:235-244: truncate or zero-extend the last dim."""
    if length <= x.shape[-1]:
        return x[..., :length]
    return F.pad(x, (0, length - x.shape[-1]))


# ----------------------------------------------------------------------------------------
# filterbanks + spectrogram front-ends (modules.py:119-214; torchaudio.functional semantics)
# ----------------------------------------------------------------------------------------
def _tri_fbanks(all_freqs: Tensor, f_pts: Tensor) -> Tensor:
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)        # (n_freqs, n+2)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def mel_fbanks(n_freqs: int, n_mels: int, sample_rate: int) -> Tensor:
    """torchaudio.functional.melscale_fbanks(htk, norm=None, f_min=0, f_max=sr/2) -> (n_freqs, n_mels)"""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (float(sample_rate // 2)) / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    return _tri_fbanks(all_freqs, f_pts)


def linear_fbanks(n_freqs: int, n_filter: int, sample_rate: int) -> Tensor:
    """torchaudio.functional.linear_fbanks(f_min=0, f_max=sr//2) -> (n_freqs, n_filter)"""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    f_pts = torch.linspace(0.0, float(sample_rate // 2), n_filter + 2)
    return _tri_fbanks(all_freqs, f_pts)


def mel_spectrogram(x: Tensor, n_fft: int, hop: int, fb: Tensor) -> Tensor:
    """torchaudio MelSpectrogram(power=1): fb^T |STFT| -> (..., n_mels, F)"""
    lead = x.shape[:-1]
    mag = stft_complex(x.reshape(-1, x.shape[-1]), n_fft, hop).abs()      # (B, nb, F)
    mel = torch.matmul(mag.transpose(1, 2), fb).transpose(1, 2)
    return mel.reshape(*lead, mel.shape[-2], mel.shape[-1])


def safe_log(x: Tensor, clip: float = 1e-7) -> Tensor:
    return torch.log(torch.clamp(x, min=clip))                            # utils.py:221-232


def log_mel(x: Tensor, sampling_rate: int = 24000, n_fft: int = 1024, hop: int = 256,
            n_mels: int = 100) -> Tensor:
    """LogMelSpectrogram.forward (modules.py:140-143)."""
    fb = mel_fbanks(n_fft // 2 + 1, n_mels, sampling_rate)
    return safe_log(mel_spectrogram(x, n_fft, hop, fb))


def linear_filter_spectrogram(x: Tensor, n_fft: int, hop: int, fb: Tensor) -> Tensor:
    """LinearFilterSpectrogram.forward, power=2 (modules.py:203-214)."""
    p = stft_complex(x, n_fft, hop).abs().pow(2.0)
    return torch.matmul(p.transpose(1, 2), fb).transpose(1, 2)


# ----------------------------------------------------------------------------------------
# generator blocks (modules.py:217-720)
# ----------------------------------------------------------------------------------------
class _LimitParam(torch.autograd.Function):
    """LimitParamValue (modules.py:236-256): identity forward; backward flips the sign of
    gradient entries that would push an out-of-range parameter further out of range."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.save_for_backward(x)
        ctx.lo, ctx.hi = lo, hi
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g * torch.where((g > 0) & (x < ctx.lo), -1.0, 1.0)
        g = g * torch.where((g < 0) & (x > ctx.hi), -1.0, 1.0)
        return g, None, None


def limit_param(x: Tensor, lo: float, hi: float, apply: bool) -> Tensor:
    """limit_param_value (modules.py:259-270) with the random draw made explicit."""
    return _LimitParam.apply(x, lo, hi) if apply else x


def bias_norm(x: Tensor, bias: Tensor, log_scale: Tensor, limit: bool = False) -> Tensor:
    """modules.py:309-312, 407-416; channel dim 1, no epsilon."""
    log_scale = limit_param(log_scale, -1.5, 1.5, limit)
    scales = torch.mean((x - bias[None, :, None]) ** 2, dim=1, keepdim=True) ** -0.5
    return x * scales * log_scale.exp()


def sinusoidal_pos_emb(t: Tensor, dim: int = 512, scale: float = 1000.0) -> Tensor:
    """modules.py:223-232."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half).float() * -e)
    a = scale * t.reshape(-1, 1) * freq[None]
    return torch.cat([a.sin(), a.cos()], dim=-1)


def convnext_block(sd: SD, pre: str, x: Tensor, cond: Optional[Tensor] = None,
                   te: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                   limit: bool = False) -> Tensor:
    """ConvNeXtBlock.forward (modules.py:456-495)."""
    r = x
    if mask is not None:
        x = x * mask
    C = x.shape[1]
    x = F.conv1d(x, sd[pre + "dwconv.weight"], sd[pre + "dwconv.bias"], padding=3, groups=C)
    x = bias_norm(x, sd[pre + "norm.bias"], sd[pre + "norm.log_scale"], limit)
    if cond is not None:
        x = x + F.conv1d(cond, sd[pre + "cond_proj.weight"], sd[pre + "cond_proj.bias"])
    if te is not None:
        x = x * (1.0 + F.linear(te, sd[pre + "time_embed_proj.weight"],
                                sd[pre + "time_embed_proj.bias"]).unsqueeze(-1))
    x = F.conv1d(x, sd[pre + "pwconv1.weight"], sd[pre + "pwconv1.bias"])
    x = F.prelu(x, sd[pre + "act.weight"])
    x = F.conv1d(x, sd[pre + "pwconv2.weight"], sd[pre + "pwconv2.bias"])
    return x + r * limit_param(sd[pre + "residual_scale.scale"], 0.5, 1.0, limit)


def cond_encoder(sd: SD, mel: Tensor, pre: str = "cond_encoder.", n_layers: int = 4,
                 limit: bool = False) -> Tensor:
    """CondEncoder.forward (modules.py:523-542)."""
    x = F.conv1d(mel, sd[pre + "in_proj.weight"], sd[pre + "in_proj.bias"], padding=1)
    x = bias_norm(x, sd[pre + "in_norm.bias"], sd[pre + "in_norm.log_scale"], limit)
    for i in range(n_layers):
        x = convnext_block(sd, f"{pre}blocks.{i}.", x, limit=limit)
    return x


def upsample_cond(cond: Tensor, factor: int, frames: int) -> Tensor:
    """AudioConvNeXt.upsample_cond (modules.py:668-680)."""
    if factor != 1:
        cond = cond.repeat_interleave(factor, dim=2)
    return convert_length(cond, frames)


def decoder(sd: SD, pre: str, x: Tensor, cond: Tensor, t: Tensor,
            mask: Optional[Tensor], n_layers: int = 8, limit: bool = False) -> Tensor:
    """ConvNeXtDecoder.forward (modules.py:595-627)."""
    x = F.conv1d(x, sd[pre + "in_proj.weight"], sd[pre + "in_proj.bias"])
    x = bias_norm(x, sd[pre + "in_norm.bias"], sd[pre + "in_norm.log_scale"], limit)
    te = sinusoidal_pos_emb(t, sd[pre + "time_mlp.0.weight"].shape[1])
    te = F.linear(te, sd[pre + "time_mlp.0.weight"], sd[pre + "time_mlp.0.bias"])
    te = F.silu(te)
    te = F.linear(te, sd[pre + "time_mlp.2.weight"], sd[pre + "time_mlp.2.bias"])
    c = F.conv1d(cond, sd[pre + "cond_mlp.0.weight"], sd[pre + "cond_mlp.0.bias"])
    c = F.prelu(c, sd[pre + "cond_mlp.1.weight"])
    c = F.conv1d(c, sd[pre + "cond_mlp.2.weight"], sd[pre + "cond_mlp.2.bias"])
    for i in range(n_layers):
        x = convnext_block(sd, f"{pre}blocks.{i}.", x, cond=c, te=te, mask=mask, limit=limit)
    return F.conv1d(x, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def audio_convnext(sd: SD, pre: str, audio: Tensor, cond: Tensor, t: Tensor,
                   n_fft: int, hop: int, cond_hop: int,
                   audio_lens: Optional[Tensor] = None, limit: bool = False) -> Tensor:
    """AudioConvNeXt.forward (modules.py:682-721)."""
    T = audio.shape[-1]
    p = stft_packed(audio, n_fft, hop)
    Fr = p.shape[-1]
    c = upsample_cond(cond, cond_hop // hop, Fr)
    mask = None
    if audio_lens is not None:
        lens = 1 + torch.div(audio_lens, hop, rounding_mode="floor")
        mask = (torch.arange(Fr)[None, :] < lens[:, None]).unsqueeze(1).float()
    y = decoder(sd, pre + "decoder.", p, c, t, mask, limit=limit)
    if mask is not None:
        y = y * mask
    return convert_length(istft_packed(y, n_fft, hop), T)


def process_model(sd: SD, cfg: dict, x: Tensor, cond: Tensor, t: Tensor,
                  audio_lens: Optional[Tensor] = None,
                  branch_weight: Optional[Tensor] = None, limit: bool = False) -> Tensor:
    """BaseAudioGenerator.process_model (generator.py:129-170).  ``branch_weight`` (B,3) is
    the (already drawn) branch-dropout weight; None == eval mode."""
    outs = []
    for i, (n, h) in enumerate(zip(cfg["n_ffts"], cfg["hop_lengths"])):
        outs.append(audio_convnext(sd, f"estimators.{i}.", x, cond, t.flatten(), n, h,
                                   cfg["mel_hop_length"], audio_lens, limit))
    o = torch.stack(outs, dim=1)
    if branch_weight is not None:
        o = o * branch_weight.unsqueeze(-1)
    return o.mean(dim=1)


def euler_infer(sd: SD, cfg: dict, noise: Tensor, cond: Tensor,
                audio_lens: Optional[Tensor] = None, n_timesteps: int = 1,
                clamp_pred: bool = False, limit: bool = False) -> Tensor:
    """BaseAudioGenerator.infer (generator.py:236-271)."""
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    t, dt = t_span[0], t_span[1] - t_span[0]
    x = noise
    for step in range(1, n_timesteps + 1):
        pred = process_model(sd, cfg, x, cond, t[None, None].expand(noise.shape[0], 1), audio_lens,
                             limit=limit)
        vt = (pred - x) / (1 - t)
        x = x + vt * dt
        t = t_span[step]
    return x.clamp(-1.0, 1.0) if clamp_pred else x


def generator_infer(sd: SD, cfg: dict, mel: Tensor, noise: Tensor,
                    audio_lens: Optional[Tensor] = None, n_timesteps: int = 1,
                    clamp_pred: bool = False, limit: bool = False) -> Tensor:
    """MelAudioGenerator.infer with the noise made explicit (generator.py:327-366).
    ``limit`` = the (pinned) limit_param_value draw when the generator is in train mode."""
    cond = cond_encoder(sd, mel, n_layers=cfg["cond_enc_num_layers"], limit=limit)
    return euler_infer(sd, cfg, noise, cond, audio_lens, n_timesteps, clamp_pred, limit)


def fm_loss(sd: SD, cfg: dict, mel: Tensor, audio: Tensor, audio_lens: Tensor,
            noise: Tensor, t: Tensor, branch_weight: Optional[Tensor] = None) -> Tensor:
    """MelAudioGenerator.forward with noise / t / branch-dropout draws made explicit
    (generator.py:294-325, 202-234, 172-200)."""
    cond = cond_encoder(sd, mel, n_layers=cfg["cond_enc_num_layers"])
    t = t.reshape(-1, 1)
    x = (1.0 - t) * noise + t * audio
    pred = process_model(sd, cfg, x, cond, t, audio_lens, branch_weight)
    err = pred - audio
    n, h = cfg["loss_n_fft"], cfg["loss_hop_length"]
    fb = linear_fbanks(n // 2 + 1, cfg["loss_n_filters"], cfg["sampling_rate"])
    gt_spec = linear_filter_spectrogram(audio, n, h, fb)
    err_spec = linear_filter_spectrogram(err, n, h, fb)
    lens = torch.div(audio_lens, h, rounding_mode="floor") + 1
    mask = (torch.arange(err_spec.shape[2])[None, :] < lens[:, None]).unsqueeze(1)
    scale = ((gt_spec + cfg["loss_eps"]) ** -cfg["loss_power"]).clamp(
        min=cfg["loss_scale_min"], max=cfg["loss_scale_max"])
    loss = err_spec * scale
    return (loss * mask).sum() / (mask.sum() * err_spec.shape[1])


# ----------------------------------------------------------------------------------------
# discriminators (discriminators.py) and GAN losses (gan.py:57-166)
# ----------------------------------------------------------------------------------------
MPD_PERIODS = (2, 3, 5, 7, 11)
MRD_WINDOWS = (2048, 1024, 512)
MRD_BAND_EDGES = (0.0, 0.1, 0.25, 0.5, 0.75, 1.0)


def discriminator_p(sd: SD, pre: str, x: Tensor, period: int) -> Tuple[Tensor, List[Tensor]]:
    """DiscriminatorP.forward (discriminators.py:79-107)."""
    x = x.unsqueeze(1)
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t += n_pad
    x = x.view(b, c, t // period, period)
    fmap = []
    for i in range(5):
        stride = (3, 1) if i < 4 else (1, 1)
        x = F.conv2d(x, sd[f"{pre}convs.{i}.weight"], sd[f"{pre}convs.{i}.bias"],
                     stride=stride, padding=(2, 0))
        x = F.leaky_relu(x, 0.1)
        if i > 0:
            fmap.append(x)
    x = F.conv2d(x, sd[pre + "conv_post.weight"], sd[pre + "conv_post.bias"], padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def mrd_bands(window: int) -> List[Tuple[int, int]]:
    nb = window // 2 + 1
    return [(int(MRD_BAND_EDGES[i] * nb), int(MRD_BAND_EDGES[i + 1] * nb)) for i in range(5)]


def mrd_spectrogram(x: Tensor, window: int) -> Tensor:
    """DiscriminatorR.spectrogram (discriminators.py:186-196) before band split:
    -> (B, 2, frames, freq)."""
    x = x - x.mean(dim=-1, keepdim=True)
    x = 0.8 * x / (x.abs().max(dim=-1, keepdim=True)[0] + 1e-9)
    s = stft_complex(x, window, window // 4)                       # (B, freq, frames)
    return torch.stack([s.real, s.imag], dim=1).transpose(2, 3)


def discriminator_r(sd: SD, pre: str, x: Tensor, window: int) -> Tuple[Tensor, List[Tensor]]:
    """DiscriminatorR.forward (discriminators.py:198-219)."""
    spec = mrd_spectrogram(x, window)
    fmap, outs = [], []
    for bi, (lo, hi) in enumerate(mrd_bands(window)):
        band = spec[..., lo:hi]
        for li in range(5):
            w = sd[f"{pre}band_convs.{bi}.{li}.weight"]
            bb = sd[f"{pre}band_convs.{bi}.{li}.bias"]
            if li == 0:
                band = F.conv2d(band, w, bb, stride=(1, 1), padding=(1, 4))
            elif li < 4:
                band = F.conv2d(band, w, bb, stride=(1, 2), padding=(1, 4))
            else:
                band = F.conv2d(band, w, bb, stride=(1, 1), padding=(1, 1))
            band = F.leaky_relu(band, 0.1)
            if li > 0:
                fmap.append(band)
        outs.append(band)
    x = torch.cat(outs, dim=-1)
    x = F.conv2d(x, sd[pre + "conv_post.weight"], sd[pre + "conv_post.bias"], padding=(1, 1))
    fmap.append(x)
    return x, fmap


def mpd(sd: SD, y: Tensor, pre: str = "discriminator.0.") -> Tuple[List[Tensor], List[List[Tensor]]]:
    scores, fmaps = [], []
    for i, p in enumerate(MPD_PERIODS):
        s, f = discriminator_p(sd, f"{pre}discriminators.{i}.", y, p)
        scores.append(s)
        fmaps.append(f)
    return scores, fmaps


def mrd(sd: SD, y: Tensor, pre: str = "discriminator.1.") -> Tuple[List[Tensor], List[List[Tensor]]]:
    scores, fmaps = [], []
    for i, w in enumerate(MRD_WINDOWS):
        s, f = discriminator_r(sd, f"{pre}discriminators.{i}.", y, w)
        scores.append(s)
        fmaps.append(f)
    return scores, fmaps


def hinge_d_loss(real: Sequence[Tensor], fake: Sequence[Tensor]) -> Tensor:
    loss = 0
    for r, f in zip(real, fake):
        loss = loss + torch.mean(torch.clamp(1 - r, min=0)) + torch.mean(torch.clamp(1 + f, min=0))
    return loss


def hinge_g_loss(fake: Sequence[Tensor]) -> Tensor:
    loss = 0
    for f in fake:
        loss = loss + torch.mean(torch.clamp(1 - f, min=0))
    return loss


def feature_matching_loss(fr: List[List[Tensor]], ff: List[List[Tensor]]) -> Tensor:
    loss = 0
    for a, b in zip(fr, ff):
        for r, f in zip(a, b):
            loss = loss + F.l1_loss(r.detach(), f)
    return loss


def mel_recon_loss(real: Tensor, fake: Tensor, sampling_rate: int,
                   n_ffts: Sequence[int] = (32, 64, 128, 256, 512, 1024, 2048),
                   n_mels: Sequence[int] = (5, 10, 20, 40, 80, 160, 320)) -> Tensor:
    """GAN.mel_recon_loss (gan.py:89-99); MelSpectrogram(hop=n/4, power=1)."""
    loss = 0
    for n, m in zip(n_ffts, n_mels):
        fb = mel_fbanks(n // 2 + 1, m, sampling_rate)
        a = safe_log(mel_spectrogram(real, n, n // 4, fb))
        b = safe_log(mel_spectrogram(fake, n, n // 4, fb))
        loss = loss + F.l1_loss(a, b)
    return loss


def gan_forward(sd: SD, cfg: dict, mel: Tensor, audio: Tensor, noise: Tensor,
                audio_lens: Optional[Tensor] = None, n_timesteps: int = 1,
                train_disc: bool = True, gen_prefix: str = "generator.", limit: bool = False):
    """GAN.forward (gan.py:101-166) with the noise draw made explicit.  ``sd`` is a GAN
    state_dict (generator.* / discriminator.*).  Returns the reference's loss tuples."""
    gsd = {k[len(gen_prefix):]: v for k, v in sd.items() if k.startswith(gen_prefix)}
    if train_disc:
        with torch.no_grad():
            fake = generator_infer(gsd, cfg, mel, noise, audio_lens, n_timesteps, False)
        sr_mp, _ = mpd(sd, audio)
        sf_mp, _ = mpd(sd, fake)
        sr_mr, _ = mrd(sd, audio)
        sf_mr, _ = mrd(sd, fake)
        return hinge_d_loss(sr_mp, sf_mp), hinge_d_loss(sr_mr, sf_mr)
    fake = generator_infer(gsd, cfg, mel, noise, audio_lens, n_timesteps, False, limit)
    _, fr_mp = mpd(sd, audio)
    sf_mp, ff_mp = mpd(sd, fake)
    _, fr_mr = mrd(sd, audio)
    sf_mr, ff_mr = mrd(sd, fake)
    return (hinge_g_loss(sf_mp), hinge_g_loss(sf_mr),
            feature_matching_loss(fr_mp, ff_mp), feature_matching_loss(fr_mr, ff_mr),
            mel_recon_loss(audio, fake, cfg["sampling_rate"]))


# ----------------------------------------------------------------------------------------
# ScaledAdam (optim.py:125-619) restated per *batch of same-shape tensors*
# ----------------------------------------------------------------------------------------
class ScaledAdamOracle:
    """Functional restatement of ScaledAdam.step for one param group (optim.py:451-507).
    Operates on a dict name->tensor (params) and name->grad; keeps its own state.  Batching
    by (dtype, shape) follows BatchedOptimizer.batched_params (optim.py:77-116) because the
    clipping norm and the scalar/non-scalar distinction depend on the *stacked* shape."""

    def __init__(self, names: Sequence[str], params: Sequence[Tensor], lr: float,
                 clipping_scale: Optional[float] = None, betas=(0.9, 0.98),
                 scalar_lr_scale=0.1, eps=1e-8, param_min_rms=1e-5, param_max_rms=3.0,
                 scalar_max=10.0, size_update_period=4, clipping_update_period=100):
        self.g = dict(lr=lr, clipping_scale=clipping_scale, betas=betas,
                      scalar_lr_scale=scalar_lr_scale, eps=eps, param_min_rms=param_min_rms,
                      param_max_rms=param_max_rms, scalar_max=scalar_max,
                      size_update_period=size_update_period,
                      clipping_update_period=clipping_update_period)
        groups: Dict[tuple, List[int]] = {}
        for i, p in enumerate(params):
            groups.setdefault((str(p.dtype), *p.shape), []).append(i)
        keys = sorted(groups.keys(), key=lambda k: [names[i] for i in groups[k]])
        self.batches = [groups[k] for k in keys]
        self.names = list(names)
        self.state: List[dict] = [dict() for _ in self.batches]

    def _clip_scale(self, stacked):
        g = self.g
        st0 = self.state[0]
        step = st0["step"]
        if g["clipping_scale"] is None or step == 0:
            return 1.0
        tot = torch.tensor(0.0)
        for (p, grad), st in zip(stacked, self.state):
            if p.numel() == p.shape[0]:
                tot = tot + (grad ** 2).sum() * (g["scalar_lr_scale"] ** 2)
            else:
                tot = tot + ((grad * st["param_rms"]) ** 2).sum()
        norm = tot.sqrt()
        period = g["clipping_update_period"]
        if "model_norms" not in st0:
            st0["model_norms"] = torch.zeros(period)
        st0["model_norms"][step % period] = norm
        irregular = [i for i in (10, 20, 40) if i < period]
        if step % period == 0 or step in irregular:
            sn = st0["model_norms"].sort()[0]
            if step in irregular:
                sn = sn[-step:]
            n = sn.numel()
            median = sn[min(n - 1, (n // 4) * 2)].item()
            thr = g["clipping_scale"] * median
            if step in irregular:
                thr *= 2.0
            st0["model_norm_threshold"] = thr
        if "model_norm_threshold" not in st0:
            return 1.0
        ans = min(1.0, (st0["model_norm_threshold"] / (norm + 1e-20)).item())
        if ans != ans:
            ans = 0.0
        return ans

    @torch.no_grad()
    def step(self, params: Sequence[Tensor], grads: Sequence[Optional[Tensor]]):
        g = self.g
        stacked = []
        for idxs in self.batches:
            p = torch.stack([params[i] for i in idxs])
            gr = torch.stack([torch.zeros_like(params[i]) if grads[i] is None else grads[i]
                              for i in idxs])
            stacked.append((p, gr))
        clip = 1.0 if len(self.state[0]) == 0 else self._clip_scale(stacked)
        if clip == 0.0:
            stacked = [(p, torch.zeros_like(gr)) for p, gr in stacked]
        for (p, grad), st, idxs in zip(stacked, self.state, self.batches):
            step = st.setdefault("step", 0)
            if clip != 1.0:
                grad = grad * clip
            delta = self._momentum_step(p, st, grad)
            p = p + delta
            if p.numel() == p.shape[0]:
                p = p.clamp(-g["scalar_max"], g["scalar_max"])
            st["step"] = step + 1
            for j, i in enumerate(idxs):
                params[i].copy_(p[j])

    def _basic_step(self, p, st, grad):
        g = self.g
        lr = g["lr"] * (g["scalar_lr_scale"] if p.numel() == p.shape[0] else 1.0)
        beta2 = g["betas"][1]
        if "exp_avg_sq" not in st:
            st["exp_avg_sq"] = torch.zeros_like(p)
        v = st["exp_avg_sq"]
        v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        bc2 = 1 - beta2 ** (st["step"] + 1)
        if bc2 < 0.99:
            v = v * (1.0 / bc2)
        return -lr * grad / (v.sqrt() + g["eps"])

    def _scaling_step(self, p, st, grad):
        g = self.g
        delta = self._basic_step(p, st, grad)
        if p.numel() == p.shape[0]:
            return delta
        step, period = st["step"], g["size_update_period"]
        dims = list(range(1, p.ndim))
        if "param_rms" not in st:
            st["param_rms"] = (p ** 2).mean(dim=dims, keepdim=True).sqrt()
            st["scale_exp_avg_sq"] = torch.zeros_like(st["param_rms"])
            st["scale_grads"] = torch.zeros(period, *st["param_rms"].shape)
        st["scale_grads"][step % period] = (p * grad).sum(dim=dims, keepdim=True)
        if step % period == period - 1:
            st["param_rms"].copy_((p ** 2).mean(dim=dims, keepdim=True).sqrt())
        rms = st["param_rms"]
        delta = delta * rms.clamp(min=g["param_min_rms"])
        if step % period == period - 1 and step > 0:
            beta2c = g["betas"][1] ** period
            size_lr = g["lr"] * g["scalar_lr_scale"]
            sg = st["scale_grads"]
            st["scale_exp_avg_sq"].mul_(beta2c).add_((sg ** 2).mean(dim=0), alpha=1 - beta2c)
            size_step = (step + 1) // period
            bc2 = 1 - beta2c ** size_step
            denom = st["scale_exp_avg_sq"].sqrt() + g["eps"]
            ss = -size_lr * (bc2 ** 0.5) * sg.sum(dim=0) / denom
            ss = ss.masked_fill(rms < g["param_min_rms"], 0.0).clamp(-0.1, 0.1)
            ss = torch.minimum(ss, (g["param_max_rms"] - rms) / rms)
            delta = delta + p * ss
        return delta

    def _momentum_step(self, p, st, grad):
        delta = self._scaling_step(p, st, grad)
        if "delta" not in st:
            st["delta"] = torch.zeros_like(p)
        st["delta"].mul_(self.g["betas"][0]).add_(delta, alpha=1 - self.g["betas"][0])
        return st["delta"]


def eden2_lr(base_lr: float, batch: int, lr_batches: float, warmup_batches: float = 500.0,
             warmup_start: float = 0.5) -> float:
    """Eden2.get_lr (optim.py:939-951)."""
    factor = ((batch ** 2 + lr_batches ** 2) / lr_batches ** 2) ** -0.5
    warm = 1.0 if batch >= warmup_batches else (
        warmup_start + (1.0 - warmup_start) * (batch / warmup_batches))
    return base_lr * factor * warm
