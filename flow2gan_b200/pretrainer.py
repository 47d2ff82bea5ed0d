"""Stage-1 (flow-matching) pre-training iteration of flow2gan/bin/pretrain.py as a small class
(SURVEY.md section 8(f).3): LogMel front-end -> `model(cond, audio, audio_lens)` -> backward ->
gradient averaging across data-parallel ranks -> Eden2.step_batch -> ScaledAdam.step, with the
optimiser built from `get_parameter_groups_with_lrs` (pretrain.py:794-799) and the fp64 running
model average updated every `average_period` batches on rank 0 (pretrain.py:477-489).

Reference step body: pretrain.py:340-359 (compute_loss) and :446-475."""
from __future__ import annotations

import copy
from typing import Dict, Optional

import torch
from torch import Tensor, nn

from .averaging import update_averaged_model
from .dist import GradBuckets, broadcast_module_state, get_rank
from .modules import LogMelSpectrogram
from .optim import Eden2, ScaledAdam
from .utils import get_parameter_groups_with_lrs


class FMTrainer:
    def __init__(self, model: nn.Module, base_lr: float = 0.035, lr_batches: float = 7500,
                 warmup_start: float = 0.1, average_period: int = 200, keep_average: bool = True,
                 rank: Optional[int] = None):
        self.model = model
        broadcast_module_state(model)            # DDP's construction-time sync (pretrain.py:800-802)
        dev = next(model.parameters()).device
        self.cond_module = LogMelSpectrogram(model.sampling_rate, model.mel_n_fft, model.mel_hop_length,
                                             model.n_mels).to(dev)
        self.optimizer = ScaledAdam(get_parameter_groups_with_lrs(model, lr=base_lr, include_names=True),
                                    lr=base_lr, clipping_scale=2.0)
        self.scheduler = Eden2(self.optimizer, lr_batches, warmup_start=warmup_start)
        self.buckets = GradBuckets(model.parameters())
        self.average_period = average_period
        self.batch_idx_train = 0
        self.rank = get_rank() if rank is None else rank
        # model_avg lives on rank 0 only, in fp64 (pretrain.py:774-777)
        self.model_avg = copy.deepcopy(model).to(torch.float64) if keep_average and self.rank == 0 else None

    def step(self, audio: Tensor, audio_lens: Tensor) -> Dict[str, Tensor]:
        """One batch: (B, T) fp32 audio, (B,) lengths -> {"loss": 0-dim tensor (no host sync)}."""
        self.model.train()
        self.batch_idx_train += 1
        with torch.no_grad():
            cond = self.cond_module(audio)
        loss = self.model(cond=cond, audio=audio, audio_lens=audio_lens)
        loss.backward()
        self.buckets.allreduce_mean()
        self.scheduler.step_batch(self.batch_idx_train)
        self.optimizer.step()
        self.optimizer.zero_grad()
        if (self.model_avg is not None and self.batch_idx_train > 0
                and self.batch_idx_train % self.average_period == 0):
            update_averaged_model({"average_period": self.average_period,
                                   "batch_idx_train": self.batch_idx_train}, self.model, self.model_avg)
        return {"loss": loss.detach()}
