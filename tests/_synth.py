"""Deterministic synthetic weights in the reference's state_dict layout -- TEST/BENCH
INFRASTRUCTURE (no reference weights can be downloaded: there is no network).

``synth_state_dict(spec, seed)`` turns a ``[(key, shape), ...]`` list into parameter tensors
whose every term is non-trivial (non-zero biases, varied PReLU slopes, residual scales and
BiasNorm log-scales) so parity tests exercise every code path.  Buffers (STFT windows,
filterbanks) are copied from ``like`` when given and skipped otherwise (each side builds its
own).  The draw depends only on (seed, key order, shapes).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch


def is_buffer_key(k: str) -> bool:
    return k.endswith("window") or k.endswith(".fb")


def synth_state_dict(spec: Sequence[Tuple[str, tuple]], seed: int = 1234,
                     like: Optional[Dict[str, torch.Tensor]] = None,
                     gain: float = 0.9, style: str = "ref_init") -> Dict[str, torch.Tensor]:
    """style="ref_init": generator matrices follow the reference's own initialisation
    (trunc_normal std=0.015, flow2gan/models/generator.py:122-127) -- the only "real" weights
    available offline -- while every small parameter is made non-trivial; discriminator
    matrices follow torch's default conv init scale (std = 0.577/sqrt(fan_in)).
    style="harsh": every matrix ~ N(0, gain^2/fan_in) (strong branches, little residual
    dilution): a stress test for accumulated TF32 rounding."""
    assert style in ("ref_init", "harsh")
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for k, shape in spec:
        shape = tuple(shape)
        if is_buffer_key(k):
            if like is not None:
                out[k] = like[k].clone()
            continue
        v = torch.empty(shape)
        if k.endswith("log_scale"):
            v.uniform_(0.4, 1.7, generator=g)      # init 1.0; BiasNorm limit range is [-1.5, 1.5]
        elif k.endswith("residual_scale.scale"):
            v.uniform_(0.45, 1.1, generator=g)     # init 1.0; ChannelScale limit range is [0.5, 1.0]
        elif k.endswith("act.weight") or k.endswith("cond_mlp.1.weight"):
            v.uniform_(0.05, 0.45, generator=g)
        elif k.endswith(".bias"):
            v.normal_(0.0, 0.05, generator=g)
        elif k.endswith(".weight"):
            fan_in = max(1, int(torch.Size(shape[1:]).numel()))
            v.normal_(0.0, 1.0, generator=g)
            if style == "harsh":
                v.mul_(gain / math.sqrt(fan_in))
            elif k.startswith("discriminator.") or "dwconv" in k:
                v.mul_(0.577 / math.sqrt(fan_in))
            else:
                v.clamp_(-2.0, 2.0).mul_(0.015)
        else:
            raise KeyError(f"synth_state_dict: unclassified key {k}")
        out[k] = v
    return out
