"""The GAN fine-tuning iteration of flow2gan/bin/finetune.py:427-492,590-626 as a small class:
alternating discriminator / generator phases, the recipe's loss weights, ScaledAdam + Eden2 per
half, gradient averaging across data-parallel ranks for the half being stepped."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from .dist import GradBuckets
from .gan import GAN
from .modules import LogMelSpectrogram
from .optim import Eden2, ScaledAdam

LOSS_WEIGHTS = dict(mp=1.0, mr=0.1, fm_mp=1.0, fm_mr=0.1, mel=45.0)      # finetune.py defaults


class GANTrainer:
    def __init__(self, gan: GAN, lr_g: float = 2e-3, lr_d: float = 2e-2, lr_batches_g: float = 20000,
                 lr_batches_d: float = 5000, warmup_start: float = 0.1, n_timesteps: int = 1,
                 weights: Optional[Dict[str, float]] = None):
        self.gan = gan
        g = gan.generator
        self.cond_module = LogMelSpectrogram(g.sampling_rate, g.mel_n_fft, g.mel_hop_length, g.n_mels) \
            .to(next(gan.parameters()).device)
        self.opt_g = ScaledAdam(gan.generator.named_parameters(), lr=lr_g, clipping_scale=2.0)
        self.opt_d = ScaledAdam(gan.discriminator.named_parameters(), lr=lr_d, clipping_scale=2.0)
        self.sched_g = Eden2(self.opt_g, lr_batches_g, warmup_start=warmup_start)
        self.sched_d = Eden2(self.opt_d, lr_batches_d, warmup_start=warmup_start)
        self.buckets_g = GradBuckets(gan.generator.parameters())
        self.buckets_d = GradBuckets(gan.discriminator.parameters())
        self.n_timesteps = n_timesteps
        self.w = dict(LOSS_WEIGHTS, **(weights or {}))
        self.train_disc = True

    def step(self, audio: Tensor, audio_lens: Tensor) -> Dict[str, Tensor]:
        """One iteration on one batch: D-phase or G-phase (they alternate, finetune.py:612-626)."""
        with torch.no_grad():
            cond = self.cond_module(audio)                                   # finetune.py:441
        w = self.w
        if self.train_disc:
            d_mp, d_mr = self.gan(cond=cond, audio=audio, audio_lens=audio_lens,
                                  n_timesteps=self.n_timesteps, train_disc=True)
            loss = d_mp * w["mp"] + d_mr * w["mr"]
            self.opt_d.zero_grad()
            loss.backward()
            self.buckets_d.allreduce_mean()
            self.opt_d.step()
            self.sched_d.step_batch()
            info = {"disc_loss": loss.detach(), "disc_loss_mp": d_mp.detach(), "disc_loss_mr": d_mr.detach()}
        else:
            g_mp, g_mr, fm_mp, fm_mr, mel = self.gan(cond=cond, audio=audio, audio_lens=audio_lens,
                                                     n_timesteps=self.n_timesteps, train_disc=False)
            loss = (g_mp * w["mp"] + g_mr * w["mr"] + fm_mp * w["fm_mp"] + fm_mr * w["fm_mr"]
                    + mel * w["mel"])
            self.opt_g.zero_grad()
            loss.backward()
            self.buckets_g.allreduce_mean()
            self.opt_g.step()
            self.sched_g.step_batch()
            info = {"gen_loss": loss.detach(), "mel_recon_loss": mel.detach()}
        self.train_disc = not self.train_disc
        return info
