"""GAN wrapper with the reference's constructor, sub-module names, forward signature, returned loss
tuples and train()/eval() side effects (flow2gan/models/gan.py:30-166)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
from torch import Tensor, nn

from .discriminators import MultiPeriodDiscriminator, MultiResolutionDiscriminator
from .losses import hinge_terms, l1_terms, mel_recon_loss
from .modules import MelSpectrogram

# Hinge / feature-matching / mel L1 reductions go through the multi-tensor kernels of csrc/losses.cu
# (one launch per <= 24 terms each way) instead of ~850 element-wise torch launches per D+G pair
# (GPU parity: tests/test_zz_losses_gpu.py; A/B 66.8 -> 65.2 ms per pair, profiles/r02_switches.md).
# F2G_FUSED_LOSSES=0 keeps the per-term torch reductions for A/B runs.
FUSED_LOSSES = os.environ.get("F2G_FUSED_LOSSES", "1") == "1"


class GAN(nn.Module):
    def __init__(self, generator: nn.Module,
                 mel_recon_n_ffts: Tuple[int, ...] = (32, 64, 128, 256, 512, 1024, 2048),
                 mel_recon_n_mels: Tuple[int, ...] = (5, 10, 20, 40, 80, 160, 320)):
        super().__init__()
        self.generator = generator
        self.discriminator = nn.ModuleList([MultiPeriodDiscriminator(), MultiResolutionDiscriminator()])
        self.mel_recon_modules = nn.ModuleList([
            MelSpectrogram(sample_rate=generator.sampling_rate, n_fft=n_fft, hop_length=n_fft // 4,
                           n_mels=n_mels, center=True, power=1)
            for n_fft, n_mels in zip(mel_recon_n_ffts, mel_recon_n_mels)])

    @staticmethod
    def discriminator_loss(score_real: List[Tensor], score_fake: List[Tensor]):
        if FUSED_LOSSES:
            return hinge_terms(list(score_real) + list(score_fake),
                               [-1.0] * len(score_real) + [1.0] * len(score_fake))
        loss = 0
        for s_real, s_fake in zip(score_real, score_fake):
            loss = loss + torch.mean(torch.clamp(1 - s_real, min=0)) + torch.mean(torch.clamp(1 + s_fake, min=0))
        return loss

    @staticmethod
    def generator_loss(score_fake: List[Tensor]):
        if FUSED_LOSSES:
            return hinge_terms(list(score_fake), [-1.0] * len(score_fake))
        loss = 0
        for s_fake in score_fake:
            loss = loss + torch.mean(torch.clamp(1 - s_fake, min=0))
        return loss

    @staticmethod
    def feature_matching_loss(fmap_real: List[List[Tensor]], fmap_fake: List[List[Tensor]]):
        if FUSED_LOSSES:
            for f_real, f_fake in zip(fmap_real, fmap_fake):
                assert isinstance(f_real, list) and isinstance(f_fake, list)
            return l1_terms([r for fr in fmap_real for r in fr], [f for ff in fmap_fake for f in ff])
        loss = 0
        for f_real, f_fake in zip(fmap_real, fmap_fake):
            assert isinstance(f_real, list) and isinstance(f_fake, list)
            for r, f in zip(f_real, f_fake):
                loss = loss + nn.functional.l1_loss(r.detach(), f)
        return loss

    def mel_recon_loss(self, real: Tensor, fake: Tensor):
        return mel_recon_loss(self.mel_recon_modules, real, fake, fused=FUSED_LOSSES)

    def forward(self, cond: Tensor, audio: Tensor, audio_lens: Optional[Tensor] = None,
                n_timesteps: int = 1, train_disc: bool = True, noise: Optional[Tensor] = None):
        """`noise` (extension) pins the generator's initial noise; default = torch global RNG."""
        mpd, mrd = self.discriminator[0], self.discriminator[1]
        if train_disc:
            self.discriminator.train()
            self.generator.eval()
            with torch.no_grad():
                pred_audio = self.generator.infer(cond=cond, audio_lens=audio_lens, n_timesteps=n_timesteps,
                                                  clamp_pred=False, noise=noise)
            score_real_mp, score_fake_mp, _, _ = mpd(y=audio, y_hat=pred_audio)
            score_real_mr, score_fake_mr, _, _ = mrd(y=audio, y_hat=pred_audio)
            return (self.discriminator_loss(score_real_mp, score_fake_mp),
                    self.discriminator_loss(score_real_mr, score_fake_mr))
        self.discriminator.eval()
        self.generator.train()
        pred_audio = self.generator.infer(cond=cond, audio_lens=audio_lens, n_timesteps=n_timesteps,
                                          clamp_pred=False, noise=noise)
        # discriminator weights are not stepped in this phase: skip their weight-gradient GEMMs
        _, score_fake_mp, fmap_real_mp, fmap_fake_mp = mpd(y=audio, y_hat=pred_audio, train_weights=False)
        _, score_fake_mr, fmap_real_mr, fmap_fake_mr = mrd(y=audio, y_hat=pred_audio, train_weights=False)
        return (self.generator_loss(score_fake_mp), self.generator_loss(score_fake_mr),
                self.feature_matching_loss(fmap_real_mp, fmap_fake_mp),
                self.feature_matching_loss(fmap_real_mr, fmap_fake_mr),
                self.mel_recon_loss(real=audio, fake=pred_audio))
