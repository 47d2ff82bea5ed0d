"""On-GPU data path either side of the generator.

Mirrors what the reference does on DataLoader worker processes / the host:

  * ``load_wav``  = ``LhotseRecordingDataset.__getitem__`` (flow2gan/dataset.py:122-175): read a
    segment, silence test, mono mix-down, sox ``norm`` gain, ``torchaudio.functional.resample``;
    also ``torchaudio.load`` + channel mean of the inference scripts (bin/infer_dir.py:217-220,
    test_from_wav.py:62-66).  Host work is file I/O and RIFF header parsing only; the payload
    segment goes to the device as raw bytes and ``f2g_pcm_decode`` / ``f2g_gain_resample`` do the rest.
  * ``pad_seq_collate`` = ``pad_seq_collate_fn`` (dataset.py:31-46): items are written straight into
    the rows of one zero-padded (B, Tmax) device buffer.
  * ``save_wav`` = ``soundfile.write(path, audio, sr)`` (bin/infer.py:208-212, bin/infer_dir.py:237):
    PCM16 conversion on the device (``f2g_pcm16_encode``), header + payload written by the host.

No CPU fallback: every array op is a kernel of csrc/datapath.cu.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from decimal import ROUND_HALF_UP, Decimal
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from . import _lib as L

_WAVE_FORMAT_PCM, _WAVE_FORMAT_IEEE_FLOAT, _WAVE_FORMAT_EXTENSIBLE = 1, 3, 0xFFFE


@dataclass
class WavInfo:
    sampling_rate: int
    channels: int
    sample_format: int        # L.PCM_S16 / PCM_S24 / PCM_S32 / PCM_F32
    bytes_per_sample: int
    data_offset: int          # byte offset of the first frame in the file
    num_frames: int

    @property
    def duration(self) -> float:
        return self.num_frames / self.sampling_rate


def parse_wav_header(buf: bytes) -> WavInfo:
    """RIFF/WAVE chunk walk (PCM 16/24/32-bit, IEEE float32, incl. WAVE_FORMAT_EXTENSIBLE)."""
    if len(buf) < 12 or buf[:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt = 12, None
    while pos + 8 <= len(buf):
        cid, size = buf[pos:pos + 4], struct.unpack_from("<I", buf, pos + 4)[0]
        body = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", buf, body)
            if tag == _WAVE_FORMAT_EXTENSIBLE and size >= 26:
                tag = struct.unpack_from("<H", buf, body + 24)[0]      # first 2 bytes of the sub-format GUID
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            if fmt is None:
                raise ValueError("wav: data chunk before fmt chunk")
            tag, ch, sr, bits = fmt
            if tag == _WAVE_FORMAT_PCM and bits in (16, 24, 32):
                sf = {16: L.PCM_S16, 24: L.PCM_S24, 32: L.PCM_S32}[bits]
            elif tag == _WAVE_FORMAT_IEEE_FLOAT and bits == 32:
                sf = L.PCM_F32
            else:
                raise ValueError(f"wav: unsupported encoding (format tag {tag}, {bits} bits)")
            bps = bits // 8
            size = min(size, len(buf) - body)              # tolerate a truncated / streamed size field
            return WavInfo(sr, ch, sf, bps, body, size // (bps * ch))
        pos = body + size + (size & 1)
    raise ValueError("wav: no data chunk")


def seconds_to_samples(seconds: float, sampling_rate: int) -> int:
    """lhotse.utils.compute_num_samples (the rounding `Recording.load_audio(offset, duration)` uses)."""
    return int(Decimal(round(seconds * sampling_rate, ndigits=8)).quantize(0, rounding=ROUND_HALF_UP))


# ---------------------------------------------------------------------------------------------
# torchaudio.functional.resample tap table (host-side constant, like a window buffer)
# ---------------------------------------------------------------------------------------------
_TAPS: Dict[Tuple[int, int, str], Tuple[Tensor, int, int, int]] = {}


def sinc_resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                       rolloff: float = 0.99) -> Tuple[Tensor, int, int, int]:
    """(taps (new_r, 2*width + orig_r) fp32 on CPU, width, orig_r, new_r): the sinc_interp_hann kernel
    torchaudio.functional.resample builds for an fp32 waveform (_get_sinc_resample_kernel with
    dtype=float32: every step below is an fp32 CPU op in the same order, so the table is the one the
    reference's DataLoader workers convolve with)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig_r, new_r = int(orig_freq) // g, int(new_freq) // g
    base = min(orig_r, new_r) * rolloff
    width = math.ceil(lowpass_filter_width * orig_r / base)
    f32 = torch.float32
    idx = torch.arange(-width, width + orig_r, dtype=f32)[None, :] / orig_r
    t = torch.arange(0, -new_r, -1, dtype=f32)[:, None] / new_r + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base / orig_r
    taps = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    taps *= window * scale
    return taps.contiguous(), width, orig_r, new_r


def _device_taps(orig_freq: int, new_freq: int, device) -> Tuple[Tensor, int, int, int]:
    dev = torch.device(device)
    key = (int(orig_freq), int(new_freq), str(dev) if dev.index is not None or dev.type != "cuda"
           else "cuda:%d" % torch.cuda.current_device())
    if key not in _TAPS:
        if orig_freq == new_freq:                 # gain-only pass: one unit tap, no neighbours
            taps, width, o, n = torch.ones(1, 1), 0, 1, 1
        else:
            taps, width, o, n = sinc_resample_taps(orig_freq, new_freq)
        _TAPS[key] = (taps.to(dev), width, o, n)
    return _TAPS[key]


def resampled_length(n_in: int, orig_freq: int, new_freq: int) -> int:
    g = math.gcd(int(orig_freq), int(new_freq))
    return int(math.ceil((new_freq // g) * n_in / (orig_freq // g)))


# ---------------------------------------------------------------------------------------------
# load
# ---------------------------------------------------------------------------------------------
def decode_segment(payload: Tensor, info: WavInfo, first_frame: int, n_frames: int,
                   out: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """payload: uint8 CUDA tensor whose byte 0 is frame `payload_first` = 0 of the segment's source
    (see load_wav).  Returns (mono (n_frames,) fp32, stats (2,) = [sum x^2, max |x|])."""
    dev = payload.device
    mono = out if out is not None else torch.empty(n_frames, device=dev, dtype=torch.float32)
    stats = torch.zeros(2, device=dev, dtype=torch.float32)
    L.pcm_decode(payload, info.sample_format, info.channels, first_frame, n_frames, mono, stats)
    return mono, stats


def gain_resample(x: Tensor, orig_freq: int, new_freq: int, stats: Optional[Tensor] = None,
                  norm_db: Optional[float] = None, out: Optional[Tensor] = None) -> Tensor:
    """sox ["norm", dB] (when norm_db is given; needs the decode statistics) + resample, one kernel."""
    n_in = x.numel()
    taps, width, o, n = _device_taps(orig_freq, new_freq, x.device)
    n_out = resampled_length(n_in, orig_freq, new_freq)
    if out is None:
        out = torch.empty(n_out, device=x.device, dtype=torch.float32)
    assert out.numel() >= n_out and out.is_contiguous()
    use_norm = norm_db is not None
    if use_norm:
        assert stats is not None, "norm gain needs the peak from decode_segment"
    L.gain_resample(x, n_in, stats if use_norm else None, norm_db if use_norm else 0.0, o, n, width, taps,
                    out, n_out)
    return out[:n_out]


@dataclass
class _Segment:
    buf: bytes
    info: WavInfo
    first: int
    n: int            # frames of the source segment
    n_out: int        # samples after resampling


def _open_segment(src: Union[str, bytes], sampling_rate: Optional[int], offset: float,
                  duration: Optional[float]) -> _Segment:
    if isinstance(src, (bytes, bytearray, memoryview)):
        buf = bytes(src)
    else:
        with open(src, "rb") as f:
            buf = f.read()
    info = parse_wav_header(buf)
    first = seconds_to_samples(offset, info.sampling_rate) if offset else 0
    first = max(0, min(first, info.num_frames))
    n = info.num_frames - first
    if duration is not None:
        n = min(n, seconds_to_samples(duration, info.sampling_rate))
    target = info.sampling_rate if sampling_rate is None else int(sampling_rate)
    return _Segment(buf, info, first, n, resampled_length(n, info.sampling_rate, target))


def _load_segment(seg: _Segment, sampling_rate: Optional[int], norm_db: Optional[float], min_rms: float,
                  dev: torch.device, out: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    info, n = seg.info, seg.n
    if n == 0:
        return torch.empty(0, device=dev), torch.ones((), device=dev, dtype=torch.bool)
    fb = info.bytes_per_sample * info.channels
    lo = info.data_offset + seg.first * fb
    # only the selected segment crosses PCIe, as raw PCM (2 B/sample for 16-bit files)
    host = torch.frombuffer(bytearray(seg.buf[lo: lo + n * fb]), dtype=torch.uint8)
    payload = host.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else host
    target = info.sampling_rate if sampling_rate is None else int(sampling_rate)
    second_pass = target != info.sampling_rate or norm_db is not None
    mono, stats = decode_segment(payload, info, 0, n, out=None if (second_pass or out is None) else out[:n])
    silence = torch.sqrt(stats[0] / n) < min_rms
    if not second_pass:
        return mono, silence
    return gain_resample(mono, info.sampling_rate, target, stats, norm_db, out=out), silence


def load_wav(src: Union[str, bytes], sampling_rate: Optional[int] = None, offset: float = 0.0,
             duration: Optional[float] = None, norm_db: Optional[float] = None, min_rms: float = 0.005,
             device: Union[str, torch.device] = "cuda", out: Optional[Tensor] = None
             ) -> Tuple[Tensor, Tensor, int]:
    """-> (audio (T,) fp32 on `device` at `sampling_rate`, silence flag (0-dim bool tensor, computed
    on the decoded segment before gain like dataset.py:129-152), source sampling rate).

    `src` is a path or the file's bytes.  `offset`/`duration` in seconds select the segment
    (rounded like lhotse); `norm_db` applies the sox `norm` effect (dataset.py:164-168: -3 in
    validation, U(-6,-1) in training); a `sampling_rate` different from the file's resamples.
    `out`: optional contiguous fp32 CUDA buffer (>= the result length) the result is written to."""
    seg = _open_segment(src, sampling_rate, offset, duration)
    y, silence = _load_segment(seg, sampling_rate, norm_db, min_rms, torch.device(device), out)
    return y, silence, seg.info.sampling_rate


def pad_seq_collate(sources: Sequence[Union[str, bytes]], sampling_rate: int, offsets: Optional[Sequence[float]] = None,
                    duration: Optional[float] = None, norm_dbs: Optional[Sequence[Optional[float]]] = None,
                    filter_silence: bool = True, min_rms: float = 0.005,
                    device: Union[str, torch.device] = "cuda") -> Tuple[Tensor, Tensor, List[int]]:
    """Loader + `pad_seq_collate_fn` (dataset.py:31-46) in one: -> (audios (B', Tmax) zero padded,
    audio_lens (B',) int32, kept source indices).  Result lengths are known from the headers, so
    every item is decoded / resampled straight into its row of the padded batch; silent items are
    dropped afterwards when filter_silence (the first item is kept if all are silent, like the
    reference)."""
    dev = torch.device(device)
    segs = [_open_segment(s, sampling_rate, offsets[i] if offsets else 0.0, duration)
            for i, s in enumerate(sources)]
    tmax = max((g.n_out for g in segs), default=0)
    audios = torch.zeros(len(segs), tmax, device=dev, dtype=torch.float32)
    flags = []
    for i, g in enumerate(segs):
        _, silence = _load_segment(g, sampling_rate, norm_dbs[i] if norm_dbs else None, min_rms, dev,
                                   audios[i] if g.n_out else None)
        flags.append(silence)
    keep = list(range(len(segs)))
    if filter_silence and segs:
        host_flags = torch.stack(flags).cpu().tolist()                     # one sync per batch
        keep = [i for i, f in enumerate(host_flags) if not f] or [0]
        if len(keep) != len(segs):
            t_keep = max(segs[i].n_out for i in keep)
            audios = audios[torch.tensor(keep, device=dev)][:, :t_keep].contiguous()
    lens = torch.tensor([segs[i].n_out for i in keep], dtype=torch.int32, device=dev)
    return audios, lens, keep


# ---------------------------------------------------------------------------------------------
# save
# ---------------------------------------------------------------------------------------------
def encode_pcm16(audio: Tensor, clamp: bool = True) -> Tensor:
    """(…,) fp32 CUDA -> int16 CUDA, libsndfile's float -> PCM_16 rule (lrintf(x * 32767))."""
    x = audio.contiguous().float()
    out = torch.empty(x.shape, device=x.device, dtype=torch.int16)
    L.pcm16_encode(x, x.numel(), clamp, out)
    return out


def wav_header_pcm16(n_frames: int, sampling_rate: int, channels: int = 1) -> bytes:
    data_bytes = n_frames * channels * 2
    return (b"RIFF" + struct.pack("<I", 36 + data_bytes) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, _WAVE_FORMAT_PCM, channels, sampling_rate, sampling_rate * channels * 2,
                        channels * 2, 16) + b"data" + struct.pack("<I", data_bytes))


def save_wav(path: str, audio: Tensor, sampling_rate: int, clamp: bool = True) -> None:
    """soundfile.write(path, audio.cpu().numpy(), sr) for a mono (T,) / (1, T) CUDA waveform."""
    a = audio.reshape(-1) if audio.dim() <= 1 or audio.shape[0] == 1 else None
    if a is None:
        raise ValueError("save_wav writes one mono waveform; got shape %s" % (tuple(audio.shape),))
    pcm = encode_pcm16(a, clamp).cpu()
    with open(path, "wb") as f:
        f.write(wav_header_pcm16(pcm.numel(), sampling_rate))
        f.write(pcm.numpy().tobytes())


# ---------------------------------------------------------------------------------------------
# dataset-shaped front (flow2gan/dataset.py:31-46,96-175) over plain wav paths
# ---------------------------------------------------------------------------------------------
class RecordingDataset:
    """`LhotseRecordingDataset` without lhotse: items are wav paths (or bytes).  `__getitem__` returns
    `(audio (T,) fp32 on the device at `sampling_rate`, silence flag, file name)` with the reference's
    segment policy -- whole file when `duration` is None, the first `duration` seconds in validation,
    a uniformly drawn segment in training, redrawn up to `max_load_times` while it is silent -- and
    its effects: sox `norm` at -3 dB (validation) or U(-1, -6) dB (training), then resampling.  The
    random draws come from numpy's global generator in the reference's order (offset(s), then gain)."""

    def __init__(self, recordings: Sequence[Union[str, bytes]], sampling_rate: int = 24000,
                 root_path: Optional[str] = None, train: bool = False, duration: Optional[float] = None,
                 apply_effects: bool = True, max_load_times: int = 1, min_rms: float = 0.005,
                 device: Union[str, torch.device] = "cuda"):
        self.recordings = list(recordings)
        self.sampling_rate = sampling_rate
        self.root_path = root_path
        self.train = train
        self.duration = duration
        self.apply_effects = apply_effects
        self.max_load_times = max_load_times
        self.min_rms = min_rms
        self.device = torch.device(device)

    def __len__(self) -> int:
        return len(self.recordings)

    def _name(self, src) -> str:
        import os
        if not isinstance(src, str):
            return "<bytes>"
        return os.path.relpath(src, self.root_path) if self.root_path is not None else src

    def __getitem__(self, index: int) -> Tuple[Tensor, Tensor, str]:
        import numpy as np
        src = self.recordings[index]
        if isinstance(src, str):
            with open(src, "rb") as f:
                src_bytes = f.read()
        else:
            src_bytes = bytes(src)
        info = parse_wav_header(src_bytes)
        offsets = [0.0]
        duration = None
        if self.duration is not None:
            duration = min(self.duration, info.duration)
            if self.train:
                offsets = []            # drawn lazily: one np.random.uniform per attempt (dataset.py:146-152)
        seg = mono = stats = silence = None
        attempts = 1 if not (self.train and self.duration is not None) else max(1, self.max_load_times)
        for k in range(attempts):
            off = offsets[k] if offsets else float(np.random.uniform(0, info.duration - duration))
            seg = _open_segment(src_bytes, None, off, duration)
            mono, silence, stats = _decode_only(seg, self.min_rms, self.device)
            if k + 1 < attempts and not bool(silence):          # host decision only when a redraw is possible
                break
        y = mono
        gain_db = None
        if self.apply_effects:
            gain_db = float(f"{np.random.uniform(-1, -6):.2f}") if self.train else -3.0    # dataset.py:165-167
        if gain_db is not None or info.sampling_rate != self.sampling_rate:
            y = gain_resample(mono, info.sampling_rate, self.sampling_rate, stats, gain_db)
        return y, silence, self._name(self.recordings[index])


def _decode_only(seg: _Segment, min_rms: float, dev: torch.device) -> Tuple[Tensor, Tensor, Tensor]:
    info, n = seg.info, seg.n
    if n == 0:
        z = torch.zeros(2, device=dev)
        return torch.empty(0, device=dev), torch.ones((), device=dev, dtype=torch.bool), z
    fb = info.bytes_per_sample * info.channels
    lo = info.data_offset + seg.first * fb
    host = torch.frombuffer(bytearray(seg.buf[lo: lo + n * fb]), dtype=torch.uint8)
    payload = host.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else host
    mono, stats = decode_segment(payload, info, 0, n)
    return mono, torch.sqrt(stats[0] / n) < min_rms, stats


def pad_seq_collate_fn(data, filter_silence: bool = True) -> Tuple[Tensor, Tensor, List[str]]:
    """dataset.py:31-46 on device items: drop silent items (keep the first if all are), zero-pad to the
    longest, int32 lengths, file names."""
    if filter_silence and data:
        flags = torch.stack([torch.as_tensor(x[1]).reshape(()).to(data[0][0].device) for x in data]).cpu().tolist()
        kept = [x for x, f in zip(data, flags) if not f] or list(data[0:1])
    else:
        kept = list(data)
    dev = kept[0][0].device
    tmax = max(x[0].numel() for x in kept)
    audios = torch.zeros(len(kept), tmax, device=dev, dtype=torch.float32)
    for r, x in enumerate(kept):
        audios[r, : x[0].numel()].copy_(x[0])
    lens = torch.tensor([x[0].numel() for x in kept], dtype=torch.int32)
    return audios, lens, [x[2] for x in kept]
