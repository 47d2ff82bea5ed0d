"""The STFT / iSTFT family (csrc/spectral.cu + csrc/fft_warp.cuh: shared-memory Stockham FFT, warp-
shuffle FFT, fused |.| / filterbank / log epilogue, overlap-add + branch mean + Euler update) on the
CPU: the same source under the cooperative host emulation (tests/_emul.py), through the product's own
wrappers and modules, against the oracle and -- for the log-mel front-end -- the reference's own
wav<->mel fixture (the only golden vectors the reference holds).
Mirrors the spectral cases of tests/test_kernels_gpu.py at sizes the emulation finishes in seconds."""
import io
import os
import wave

import numpy as np
import pytest
import torch

import _emul
from _cases import GOLDEN, audio_input, rel_rms
from oracle import flow2gan_oracle as O

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


@pytest.fixture
def L(monkeypatch):
    return _emul.native_fixture(monkeypatch)


def test_wav_to_logmel_reproduces_reference_fixture(L, monkeypatch):
    """wav bytes -> f2g_pcm_decode -> fused STFT / mel / log kernel == the reference's mel fixture
    (test_from_wav.py:62-70), every array op being the product's kernel source."""
    import flow2gan_b200.datapath as D
    from flow2gan_b200.modules import LogMelSpectrogram
    monkeypatch.setattr(D, "_TAPS", {})
    g = torch.load(os.path.join(GOLDEN, "mel_24k_short.pt"), weights_only=False)
    bio = io.BytesIO()
    with wave.open(bio, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000)
        w.writeframes(g["pcm_int16"].numpy().astype("<i2").tobytes())
    audio, _, sr = D.load_wav(bio.getvalue(), sampling_rate=24000, device="cpu")
    assert sr == 24000
    mel = LogMelSpectrogram(24000, 1024, 256, 100)(audio[None])
    assert mel.shape == g["mel"].shape
    assert rel_rms(mel, g["mel"]) < 1e-5


@pytest.mark.parametrize("n_fft,hop,T", [(128, 64, 1500), (32, 8, 300), (2048, 512, 4096)])
def test_stft_packed(L, n_fft, hop, T):
    B = 2
    x = audio_input(B, T, seed=n_fft)
    ref = O.stft_packed(x, n_fft, hop)
    F = ref.shape[-1]
    ld = n_fft + 8
    out = torch.full((B * F, ld), 7.0)
    L.stft(x, B, T, T, n_fft, hop, L.SPEC_PACKED, out, ld)
    got = out.view(B, F, ld)[:, :, : n_fft + 2].transpose(1, 2)
    assert rel_rms(got, ref) < 2e-6
    assert float(out[:, n_fft + 2:].abs().max()) == 0.0


@pytest.mark.parametrize("n_ffts,T", [((512, 256, 128), 2048), ((1024, 512, 256), 2500), ((256,), 1000)])
def test_warp_fft_group_vs_oracle(L, n_ffts, T):
    B = 2
    x = audio_input(B, T, seed=T)
    outs, probs = [], []
    for n in n_ffts:
        hop = n // 2
        F = 1 + T // hop
        ld = n + 4
        o = torch.full((B * F, ld), 7.0)
        outs.append((n, hop, F, ld, o))
        probs.append((x, o, n, hop, F, B * F, T, ld))
    L.stft_group(probs, B, T, round_tf32=0)
    frs, iprobs = [], []
    for n, hop, F, ld, o in outs:
        ref = O.stft_packed(x, n, hop)
        got = o.view(B, F, ld)[:, :, : n + 2].transpose(1, 2)
        assert rel_rms(got, ref) < 2e-6, n
        assert float(o[:, n + 2:].abs().max()) == 0.0
        fr = torch.empty(B * F, n)
        frs.append(fr)
        iprobs.append((o, fr, n, 0, 0, B * F, ld, n))
    L.irfft_group(iprobs)
    for (n, hop, F, ld, o), fr in zip(outs, frs):
        ref_fr = torch.empty(B * F, n)
        L.irfft_frames(o, B * F, ld, n, ref_fr)               # shared-memory Stockham kernel
        assert rel_rms(fr, ref_fr) < 2e-6, n
        out = torch.empty(B, T)
        L.ola_combine([fr], [n], [hop], [F], None, None, out, B, T, False, 0.0, 0.0, False)
        keep = hop * (F - 1)
        assert rel_rms(out[:, :keep], x[:, :keep]) < 3e-6      # STFT -> iSTFT round trip


def test_istft_vs_oracle_and_ola_combine_mean_euler_clamp(L):
    B, T = 2, 1536
    g = torch.Generator().manual_seed(5)
    cfgs = [(512, 256), (256, 128), (128, 64)]
    frs, outs = [], []
    for n, h in cfgs:
        F = 1 + T // h
        p = torch.randn(B, n + 2, F, generator=g) * 3
        outs.append(O.convert_length(O.istft_packed(p, n, h), T))
        rows = p.transpose(1, 2).reshape(B * F, n + 2).contiguous()
        fr = torch.empty(B * F, n)
        L.irfft_frames(rows, B * F, n + 2, n, fr)
        frs.append(fr)
        single = torch.empty(B, T)
        L.ola_combine([fr], [n], [h], [F], None, None, single, B, T, False, 0.0, 0.0, False)
        assert rel_rms(single, outs[-1]) < 3e-6
    x = torch.randn(B, T, generator=g)
    pred = torch.stack(outs, 1).mean(1)
    t, dt = 0.25, 0.25
    ref = (x + (pred - x) / (1 - torch.tensor(t)) * torch.tensor(dt)).clamp(-1, 1)
    xg = x.clone()
    L.ola_combine(frs, [c[0] for c in cfgs], [c[1] for c in cfgs], [1 + T // c[1] for c in cfgs],
                  None, xg, xg, B, T, True, t, dt, True)
    assert rel_rms(xg, ref) < 3e-6
    w = torch.rand(B, 3, generator=g)
    out = torch.empty(B, T)
    L.ola_combine(frs, [c[0] for c in cfgs], [c[1] for c in cfgs], [1 + T // c[1] for c in cfgs],
                  w, None, out, B, T, False, 0.0, 0.0, False)
    assert rel_rms(out, (torch.stack(outs, 1) * w[:, :, None]).sum(1)) < 3e-6


def test_logmel_44k_and_dc_peak(L):
    from flow2gan_b200.modules import LogMelSpectrogram
    x = audio_input(1, 6000, seed=9)
    m44 = LogMelSpectrogram(44100, 2048, 512, 128)
    assert rel_rms(m44(x), O.log_mel(x, 44100, 2048, 512, 128)) < 1e-5
    a = audio_input(3, 3000, seed=3) + 0.1
    pre = torch.empty(3, 2)
    L.dc_peak(a, 3, 3000, 3000, pre)
    mean = a.mean(-1)
    sc = 0.8 / ((a - mean[:, None]).abs().max(-1)[0] + 1e-9)
    assert rel_rms(pre, torch.stack([mean, sc], 1)) < 1e-5


def test_too_short_signals_are_refused_like_torch_stft(L):
    """torch.stft(center=True, pad_mode="reflect") raises when n_fft // 2 >= T; the kernels would
    otherwise reflect out of bounds."""
    x = torch.zeros(1, 256)
    with pytest.raises(RuntimeError, match="signal too short"):
        L.stft(x, 1, 256, 256, 512, 128, L.SPEC_PACKED, torch.empty(3, 520), 520)
    with pytest.raises(RuntimeError, match="signal too short"):
        L.stft_group([(x, torch.empty(5, 516), 512, 256, 2, 2, 256, 516)], 1, 256, round_tf32=0)
    with pytest.raises(RuntimeError, match="power of two"):
        L.stft(torch.zeros(1, 4000), 1, 4000, 4000, 384, 96, L.SPEC_PACKED, torch.empty(42, 392), 392)
    with pytest.raises(RuntimeError):
        torch.stft(x, 512, 128, window=torch.hann_window(512), center=True, pad_mode="reflect", return_complex=True)
