"""Seeded input builders shared by tests/golden/make_golden.py, the tests and bench.py.

Everything here is plain torch on CPU; inputs follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def mel_input(b: int, n_mels: int, frames: int, seed: int = 0, cfg44: bool = False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n_mels, frames, generator=g)
    return x * 1.4 + 0.2 if cfg44 else x * 1.7 - 1.6


def noise_input(b: int, t: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, t, generator=g) * 0.1


def audio_input(b: int, t: int, seed: int = 2):
    """A non-white synthetic 'speech-like' signal: a few chirps + noise, in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    n = torch.arange(t, dtype=torch.float32)[None]
    f0 = torch.rand(b, 1, generator=g) * 0.02 + 0.005
    sig = 0.3 * torch.sin(2 * torch.pi * f0 * n * (1 + 0.3 * n / t))
    sig = sig + 0.15 * torch.sin(2 * torch.pi * (f0 * 7.3) * n)
    sig = sig + torch.randn(b, t, generator=g) * 0.05
    return sig.clamp(-1, 1)


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b||  (the tolerance metric named in BASELINE.json: 1e-3 rel-RMS)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))
