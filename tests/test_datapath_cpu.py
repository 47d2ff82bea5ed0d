"""CPU checks of the data path:
  (1) oracle/datapath_oracle.py against the reference-generated vectors tests/golden/ref_datapath.pt
      (and against torchaudio / the mounted reference when they are importable);
  (2) the kernels of csrc/datapath.cu, compiled for the host from the same source (tests/_emul.py),
      against that oracle -- index arithmetic, padding, rounding, chunk tables;
  (3) the host logic (RIFF parsing, tap tables, parameter groups)."""
import io
import math
import os
import struct
import sys
import wave

import numpy as np
import pytest
import torch

import _emul
from _cases import GOLDEN
from oracle import datapath_oracle as DO

sys.path.insert(0, GOLDEN)
G = torch.load(os.path.join(GOLDEN, "ref_datapath.pt"), weights_only=False)
needs_gxx = pytest.mark.skipif(not _emul.available(), reason="g++ not available")
AVG_TAGS = ["running_fp32", "ema_fp32", "interval_fp64", "running_fp32_acc32", "interval_fp64_acc32"]


# ------------------------------------------------------------------------------------ oracle
@pytest.mark.parametrize("case", G["resample"], ids=lambda c: f"{c['orig']}-{c['new']}-{c['x'].numel()}")
def test_oracle_resample_matches_torchaudio_golden(case):
    y = DO.resample(case["x"][None], case["orig"], case["new"])[0]
    assert y.shape == case["y"].shape
    # conv1d's fp32 summation order depends on the CPU backend's blocking / thread split
    assert torch.allclose(y, case["y"], rtol=0, atol=1e-6), float((y - case["y"]).abs().max())


def test_oracle_resample_matches_torchaudio_live():
    AF = pytest.importorskip("torchaudio.functional")
    x = torch.randn(2, 3001, generator=torch.Generator().manual_seed(3)) * 0.2
    for o, n in ((44100, 24000), (32000, 24000), (24000, 24000)):
        assert torch.allclose(DO.resample(x, o, n), AF.resample(x, orig_freq=o, new_freq=n), rtol=0, atol=1e-6)


def _avg_case(tag):
    from make_golden_datapath import avg_inputs, clone_sd, to_f32
    avg, cur32, cur64 = avg_inputs()
    g = next(a for a in G["avg"] if a["tag"] == tag)
    return (to_f32(avg) if tag.endswith("_acc32") else clone_sd(avg)), (cur64 if "fp64" in tag else cur32), g


@pytest.mark.parametrize("tag", AVG_TAGS)
def test_oracle_average_state_dict_matches_reference(tag):
    avg, cur, g = _avg_case(tag)
    DO.average_state_dict(avg, cur, g["w1"], g["w2"], g["scale"])
    for k, v in g["result"].items():
        assert torch.equal(avg[k], v), k


@pytest.mark.parametrize("case", G["param_groups"], ids=lambda c: c["tag"])
def test_parameter_groups_match_reference(case):
    from make_golden_datapath import toy_model
    from flow2gan_b200.utils import get_parameter_groups_with_lrs
    m = toy_model(case["ddp"])
    want = [(lr, list(names)) for lr, names in case["groups"]]
    assert DO.parameter_groups(m, 0.035, case["freeze"]) == want
    mine = get_parameter_groups_with_lrs(m, lr=0.035, include_names=True, freeze_modules=case["freeze"])
    assert [(g["lr"], [n for n, _ in g["named_params"]]) for g in mine] == want
    plain = get_parameter_groups_with_lrs(m, lr=0.035, freeze_modules=case["freeze"])
    named = dict(m.named_parameters())
    for g, (lr, names) in zip(plain, want):
        assert g["lr"] == lr and [id(p) for p in g["params"]] == [id(named[n]) for n in names]


@pytest.mark.skipif(not os.path.isdir("/root/reference/flow2gan"), reason="reference not mounted")
def test_parameter_groups_on_generator_match_reference_in_place():
    from make_golden import import_reference
    import_reference()
    from flow2gan.utils import get_parameter_groups_with_lrs as ref_fn
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.utils import get_parameter_groups_with_lrs
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    m.estimators[1].lr_scale = 0.5
    m.cond_encoder.blocks[2].lr_scale = 0.1
    a = ref_fn(m, lr=0.035, include_names=True)
    b = get_parameter_groups_with_lrs(m, lr=0.035, include_names=True)
    assert [(g["lr"], [n for n, _ in g["named_params"]]) for g in a] == \
           [(g["lr"], [n for n, _ in g["named_params"]]) for g in b]


# ---------------------------------------------------------------------------------- host logic
def _wav_bytes_16(x_i16: np.ndarray, sr: int, channels: int) -> bytes:
    bio = io.BytesIO()
    with wave.open(bio, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(x_i16.astype("<i2").tobytes())
    return bio.getvalue()


def test_wav_header_parsing_and_writing():
    from flow2gan_b200 import _lib as L
    from flow2gan_b200.datapath import parse_wav_header, seconds_to_samples, wav_header_pcm16
    x = (np.arange(2 * 1001) % 777 - 300).astype(np.int16)
    info = parse_wav_header(_wav_bytes_16(x, 44100, 2))
    assert (info.sampling_rate, info.channels, info.sample_format, info.num_frames) == (44100, 2, L.PCM_S16, 1001)
    assert info.data_offset == 44 and info.bytes_per_sample == 2
    # our own writer is read back by the stdlib reader and by our parser
    payload = x[:500].astype("<i2").tobytes()
    blob = wav_header_pcm16(500, 24000) + payload
    with wave.open(io.BytesIO(blob), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 500)
        assert w.readframes(500) == payload
    assert parse_wav_header(blob).num_frames == 500
    # extensible float32 with a LIST chunk (odd size, padded) before the data chunk
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 48000, 48000 * 4, 4, 32) + struct.pack("<HHI", 22, 32, 4) + \
        struct.pack("<H", 3) + b"\x00" * 14
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + \
        b"data" + struct.pack("<I", 40) + b"\x00" * 40
    info = parse_wav_header(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert (info.sample_format, info.num_frames, info.sampling_rate) == (L.PCM_F32, 10, 48000)
    for bad in (b"RIFX" + b"\x00" * 40, b"RIFF\x00\x00\x00\x00WAVEdata\x04\x00\x00\x00abcd"):
        with pytest.raises(ValueError):
            parse_wav_header(bad)
    assert seconds_to_samples(1.5, 24000) == 36000 and seconds_to_samples(0.00002083, 24000) == 0
    assert seconds_to_samples(2.5 / 24000, 24000) == 3            # half rounds up (lhotse)


def test_tap_table_is_the_one_torchaudio_builds():
    from flow2gan_b200.datapath import resampled_length, sinc_resample_taps
    for o, n in ((44100, 24000), (16000, 24000), (48000, 24000), (24000, 44100)):
        taps, width, orr, nr = sinc_resample_taps(o, n)
        k, w2, o2, n2 = DO.sinc_kernel(o, n)
        assert (width, orr, nr) == (w2, o2, n2) and torch.equal(taps, k[:, 0])
        for length in (1, 7, 1000, 12345):
            assert resampled_length(length, o, n) == DO.resample(torch.zeros(1, length), o, n).shape[-1]


# ------------------------------------------------------------ kernels, host-emulated (same source)
def _pcm_cases():
    rng = np.random.default_rng(0)
    i16 = rng.integers(-32768, 32768, size=3 * 1000, dtype=np.int64).astype("<i2")
    i16[:6] = [-32768, 32767, 0, -1, 1, 12345]
    i24 = rng.integers(-(1 << 23), 1 << 23, size=2 * 700, dtype=np.int64)
    i24[:4] = [-(1 << 23), (1 << 23) - 1, -1, 0]
    b24 = np.stack([(i24 >> s) & 0xFF for s in (0, 8, 16)], axis=1).astype(np.uint8).tobytes()
    i32 = rng.integers(-(1 << 31), 1 << 31, size=900, dtype=np.int64).astype("<i4")
    f32 = rng.standard_normal(4 * 300).astype("<f4")
    return [("s16x1", i16.tobytes(), 16, 1), ("s16x2", i16.tobytes(), 16, 2), ("s16x3", i16.tobytes(), 16, 3),
            ("s24x2", b24, 24, 2), ("s24x1", b24, 24, 1), ("s32x1", i32.tobytes(), 32, 1),
            ("f32x4", f32.tobytes(), 1, 4)]


@needs_gxx
@pytest.mark.parametrize("name,raw,fmt,ch", _pcm_cases(), ids=[c[0] for c in _pcm_cases()])
def test_emulated_pcm_decode(name, raw, fmt, ch):
    frames = len(raw) // ((fmt if fmt != 1 else 32) // 8 * ch)
    for first, n in ((0, frames), (5, frames - 5), (frames - 1, 1), (17, 256), (3, 0)):
        rc, mono, stats = _emul.pcm_decode(raw, fmt, ch, first, n)
        assert rc == 0
        want, ss, pk = DO.pcm_decode_mono(raw, fmt, ch, first, n)
        if ch in (1, 2, 4):
            assert np.array_equal(mono, want), name                   # power-of-two means are exact
        else:
            assert np.allclose(mono, want, rtol=2e-7, atol=0)
        if n:
            assert abs(stats[0] - ss) <= 2e-5 * max(ss, 1e-6) and stats[1] == np.float32(pk)
    assert _emul.pcm_decode(raw, 8, ch, 0, 1)[0] != 0 and _emul.pcm_decode(raw, fmt, 0, 0, 1)[0] != 0


@needs_gxx
def test_emulated_pcm_decode_reference_fixture():
    """PCM16 payload of the reference's wav fixture -> exactly int16 / 32768 (what torchaudio.load
    returns, test_from_wav.py:62), which tests/test_oracle_vs_golden.py turns into the fixture's mel."""
    g = torch.load(os.path.join(GOLDEN, "mel_24k_short.pt"), weights_only=False)
    pcm = g["pcm_int16"].numpy().astype("<i2")
    rc, mono, _ = _emul.pcm_decode(pcm.tobytes(), 16, 1, 0, pcm.size)
    assert rc == 0 and np.array_equal(mono, pcm.astype(np.float32) / np.float32(32768.0))


@needs_gxx
@pytest.mark.parametrize("case", G["resample"], ids=lambda c: f"{c['orig']}-{c['new']}-{c['x'].numel()}")
def test_emulated_resample_matches_torchaudio_golden(case):
    from flow2gan_b200.datapath import resampled_length, sinc_resample_taps
    taps, width, o, n = sinc_resample_taps(case["orig"], case["new"])
    x = case["x"].numpy()
    n_out = resampled_length(x.size, case["orig"], case["new"])
    assert n_out == case["y"].numel()
    rc, y = _emul.gain_resample(x, None, 0.0, o, n, width, taps.numpy(), n_out)
    assert rc == 0
    err = np.abs(y - case["y"].numpy()).max()
    assert err < 1e-6, err                        # fp32 sums in a different order than conv1d
    # one more output than the padded signal can yield is refused, not read out of bounds
    assert _emul.gain_resample(x, None, 0.0, o, n, width, taps.numpy(), (x.size // o + 1) * n + 1)[0] != 0


@needs_gxx
def test_emulated_norm_gain_then_resample():
    from flow2gan_b200.datapath import resampled_length, sinc_resample_taps
    x = (torch.randn(5000, generator=torch.Generator().manual_seed(8)) * 0.05).numpy()
    _, ss, pk = DO.pcm_decode_mono(x.astype("<f4").tobytes(), 1, 1, 0, x.size)
    stats = np.array([ss, pk], dtype=np.float32)
    for db in (-3.0, -1.0, -6.0):
        g = DO.peak_norm_gain(x, db)
        # gain only (same rate): peak lands on the target level
        rc, y = _emul.gain_resample(x, stats, db, 1, 1, 0, np.ones((1, 1), np.float32), x.size)
        assert rc == 0 and np.allclose(y, x * g, rtol=3e-7, atol=0)
        assert abs(20 * math.log10(np.abs(y).max()) - db) < 1e-4
        # gain + resample in one pass == resample(gain * x)
        taps, width, o, n = sinc_resample_taps(44100, 24000)
        n_out = resampled_length(x.size, 44100, 24000)
        rc, y = _emul.gain_resample(x, stats, db, o, n, width, taps.numpy(), n_out)
        want = DO.resample(torch.from_numpy(x * g)[None], 44100, 24000)[0].numpy()
        assert rc == 0 and np.abs(y - want).max() < 1e-6


@needs_gxx
def test_emulated_pcm16_encode_bit_exact():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 5000).astype(np.float32),
                        np.array([0.0, 1.0, -1.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.5 / 32767,
                                  -1.5 / 32767, 3.0, -3.0, 0.99999], dtype=np.float32)])
    rc, got = _emul.pcm16_encode(x, True)
    assert rc == 0 and np.array_equal(got, DO.pcm16_encode(x, True))
    assert got.max() == 32767 and got.min() == -32767
    inside = x[np.abs(x) <= 1]
    assert np.array_equal(_emul.pcm16_encode(inside, False)[1], DO.pcm16_encode(inside, False))
    # decode(encode(x)) is within half a quantisation step
    back = got[np.abs(x) <= 1].astype(np.float32) / 32767
    assert np.abs(back - inside).max() <= 0.5 / 32767 + 1e-7


@needs_gxx
@pytest.mark.parametrize("tag", AVG_TAGS)
def test_emulated_average_update_bit_exact_vs_reference(tag):
    avg, cur, g = _avg_case(tag)
    keys = [k for k in ("w1", "b1", "s", "tail")]
    pairs = [(avg[k].numpy().reshape(-1), cur[k].numpy().reshape(-1)) for k in keys]
    pairs = [(np.ascontiguousarray(a), np.ascontiguousarray(c)) for a, c in pairs]
    assert _emul.average_update(pairs, g["w1"], g["w2"], g["scale"]) == 0
    for k, (a, _) in zip(keys, pairs):
        assert np.array_equal(a, g["result"][k].numpy().reshape(-1)), k


@needs_gxx
def test_emulated_average_update_ragged_chunks():
    rng = np.random.default_rng(2)
    sizes = [1, 4095, 4096, 4097, 10000, 3, 5000, 4098]
    avg = [rng.standard_normal(s).astype(np.float32 if i >= 6 else np.float64) for i, s in enumerate(sizes)]
    cur = [rng.standard_normal(s).astype(np.float32 if i % 2 else np.float64) for i, s in enumerate(sizes)]
    sd1 = {str(i): torch.from_numpy(a.copy()) for i, a in enumerate(avg)}
    sd2 = {str(i): torch.from_numpy(c.copy()) for i, c in enumerate(cur)}
    DO.average_state_dict(sd1, sd2, 0.973, 0.027, 1.25)
    assert _emul.average_update(list(zip(avg, cur)), 0.973, 0.027, 1.25) == 0
    for i, a in enumerate(avg):
        assert np.array_equal(a, sd1[str(i)].numpy()), i
    assert _emul.average_update([], 1.0, 0.0, 1.0) != 0


# ------------------------------------------- host layer dry run: product Python + emulated kernels
@pytest.fixture
def emulated_lib(monkeypatch):
    """flow2gan_b200._lib redirected to the host-emulated build of the same kernels (tests/_emul.py), so
    that datapath.py / averaging.py run end to end on CPU tensors through the product's own wrappers.
    Nothing in the product does this: without the patch the calls below raise (no CPU fallback)."""
    L = _emul.native_fixture(monkeypatch)
    import flow2gan_b200.datapath as D
    monkeypatch.setattr(D, "_TAPS", {})
    import flow2gan_b200.averaging as A
    monkeypatch.setattr(A, "_TABLES", {})
    return L


@needs_gxx
def test_host_layer_dry_run_load_and_decode(emulated_lib):
    import _datapath_cases as DC
    DC.case_load_fixture("cpu")
    DC.case_decode_formats("cpu")


@needs_gxx
@pytest.mark.parametrize("case", G["resample"], ids=lambda c: f"{c['orig']}-{c['new']}-{c['x'].numel()}")
def test_host_layer_dry_run_resample(emulated_lib, case):
    import _datapath_cases as DC
    DC.case_resample_golden("cpu", case)


@needs_gxx
def test_host_layer_dry_run_collate_and_save(emulated_lib, tmp_path):
    import _datapath_cases as DC
    DC.case_norm_resample_collate("cpu")
    DC.case_encode_save_round_trip("cpu", tmp_path)


@needs_gxx
def test_host_layer_dry_run_recording_dataset(emulated_lib, tmp_path):
    import _datapath_cases as DC
    DC.case_recording_dataset("cpu", tmp_path)


@needs_gxx
@pytest.mark.parametrize("tag", AVG_TAGS)
def test_host_layer_dry_run_average_state_dict(emulated_lib, tag):
    import _datapath_cases as DC
    DC.case_average_state_dict("cpu", tag)


@needs_gxx
def test_host_layer_dry_run_model_average_helpers(emulated_lib, tmp_path):
    import _datapath_cases as DC
    DC.case_model_average_helpers("cpu", tmp_path)
    DC.case_average_checkpoints("cpu", tmp_path)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less box")
def test_data_path_has_no_cpu_fallback():
    from flow2gan_b200.averaging import average_state_dict
    from flow2gan_b200.datapath import encode_pcm16, load_wav
    blob = _wav_bytes_16(np.zeros(100, np.int16), 24000, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
        load_wav(blob, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
        encode_pcm16(torch.zeros(10))
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
        average_state_dict({"w": torch.zeros(3, dtype=torch.float64)}, {"w": torch.zeros(3)}, 0.5, 0.5)
