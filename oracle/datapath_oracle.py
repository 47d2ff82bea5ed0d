"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the data path either side of the generator
(SURVEY.md section 8(f) rows 3-4).  Only tests/ and __graft_entry__.smoke() may import this.

Pinning:
  * ``resample``            -- restates torchaudio.functional.resample (functional.py
    _get_sinc_resample_kernel / _apply_sinc_resample_kernel; un-vendored dependency, unpinned in
    requirements.txt:8; call site flow2gan/dataset.py:170-173).  Pinned against torchaudio itself
    (2.11.0, importable in the build container) and the committed vectors tests/golden/ref_datapath.pt.
  * ``pcm_decode_mono``     -- libsndfile / torchaudio.load integer scaling (2^-(bits-1)) + channel
    mean (dataset.py:136-160, infer_dir.py:217-220).  Pinned by the reference's wav<->mel fixtures
    (PCM16: tests/golden/mel_24k_short.pt; the stereo 44.1 kHz pair in place when /root/reference is
    mounted).  24/32-bit/float paths: parity unpinned (no fixture exists), restated from the format.
  * ``peak_norm_gain``      -- sox effect ["norm", dB] (dataset.py:164-168): sox is absent from this
    image (and from torchaudio 2.11), so this restates the published behaviour of `gain -n dB`
    (scale so that the peak magnitude sits at dB FS).  **parity unpinned**.
  * ``pcm16_encode``        -- soundfile.write default subtype PCM_16 = libsndfile f2s_array,
    lrintf(x * 0x7FFF) (infer.py:212, infer_dir.py:237).  soundfile is absent: **parity unpinned**.
  * ``average_state_dict``  -- flow2gan/checkpoint.py:504-531 restated op for op; pinned against the
    imported reference function in tests/golden/make_golden_datapath.py.
  * ``parameter_groups``    -- flow2gan/utils.py:69-138; pinned the same way.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------- decode
def pcm_decode_mono(raw: bytes, sample_format: int, channels: int, first: int, n: int
                    ) -> Tuple[np.ndarray, float, float]:
    """-> (mono float32 (n,), sum of squares, peak magnitude) of frames [first, first + n)."""
    if sample_format == 16:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif sample_format == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) * np.float32(1.0 / 8388608.0)
    elif sample_format == 32:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) * np.float32(1.0 / 2147483648.0)
    elif sample_format == 1:
        x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    else:
        raise ValueError(sample_format)
    x = x.reshape(-1, channels)[first:first + n]
    if channels > 1:
        s = x[:, 0].copy()
        for c in range(1, channels):
            s = (s + x[:, c]).astype(np.float32)
        mono = (s / np.float32(channels)).astype(np.float32)
    else:
        mono = x[:, 0].copy()
    ss = float(np.sum(mono.astype(np.float64) ** 2))
    pk = float(np.max(np.abs(mono))) if n else 0.0
    return mono, ss, pk


def is_silence(mono: np.ndarray, min_rms: float = 0.005) -> bool:
    """dataset.py:129-130"""
    return bool(np.sqrt(np.mean(mono.astype(np.float64) ** 2)) < min_rms)


def peak_norm_gain(mono: np.ndarray, db: float) -> np.float32:
    """linear gain of sox `norm dB`: 10^(dB/20) / max|x|"""
    pk = np.float32(max(float(np.max(np.abs(mono))), 1.0e-30))
    return np.float32(np.float32(10.0 ** (db / 20.0)) / pk)


# ----------------------------------------------------------------------------------- resample
def sinc_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99
                ) -> Tuple[torch.Tensor, int, int, int]:
    """torchaudio _get_sinc_resample_kernel, sinc_interp_hann, dtype=float32 (what
    functional.resample passes for an fp32 waveform)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base_freq)
    idx = torch.arange(-width, width + o, dtype=torch.float32)[None, None] / o
    t = torch.arange(0, -n, -1, dtype=torch.float32)[:, None, None] / n + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / o
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels, width, o, n


def resample(x: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """torchaudio.functional.resample for a (…, T) fp32 waveform."""
    if orig_freq == new_freq:
        return x
    kernels, width, o, n = sinc_kernel(orig_freq, new_freq)
    shape = x.shape
    w = x.reshape(-1, shape[-1])
    length = w.shape[1]
    w = F.pad(w, (width, width + o))
    y = F.conv1d(w[:, None], kernels, stride=o)
    y = y.transpose(1, 2).reshape(w.shape[0], -1)
    target = int(math.ceil(n * length / o))
    return y[..., :target].reshape(shape[:-1] + (target,))


# ------------------------------------------------------------------------------------- encode
def pcm16_encode(x: np.ndarray, clamp: bool = True) -> np.ndarray:
    v = x.astype(np.float32)
    if clamp:
        v = np.clip(v, np.float32(-1.0), np.float32(1.0))
    return np.rint(v * np.float32(32767.0)).astype(np.int16)           # rint = round-half-even = lrintf


# ---------------------------------------------------------------------------------- averaging
def average_state_dict(sd1: Dict[str, torch.Tensor], sd2: Dict[str, torch.Tensor], weight_1: float,
                       weight_2: float, scaling_factor: float = 1.0) -> Dict[str, torch.Tensor]:
    """checkpoint.py:504-531 (in place on sd1)."""
    uniq: "OrderedDict[int, str]" = OrderedDict()
    for k, v in sd1.items():
        if v.data_ptr() not in uniq:
            uniq[v.data_ptr()] = k
    for k in uniq.values():
        v = sd1[k]
        if torch.is_floating_point(v):
            v *= weight_1
            v += sd2[k].to(device=v.device) * weight_2
            v *= scaling_factor
    return sd1


# ------------------------------------------------------------------------------ param groups
def parameter_groups(model: torch.nn.Module, lr: float, freeze_modules=()) -> List[Tuple[float, List[str]]]:
    """utils.py:69-138 as [(lr, [parameter names])] in group order."""
    scale = {name: float(m.lr_scale) for name, m in model.named_modules() if hasattr(m, "lr_scale")}
    out: "OrderedDict[float, List[str]]" = OrderedDict()
    for name, _ in model.named_parameters():
        parts = name.split(".")
        head = parts[0]
        if head == "module":
            if parts[1] in freeze_modules:
                continue
        elif head in freeze_modules:
            continue
        cur = lr * scale.get(head, 1.0)
        if head != "":
            cur *= scale.get("", 1.0)
        pre = head
        for p in parts[1:]:
            pre = pre + "." + p
            cur *= scale.get(pre, 1.0)
        out.setdefault(cur, []).append(name)
    return list(out.items())
