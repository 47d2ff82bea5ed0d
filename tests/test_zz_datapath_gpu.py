"""GPU parity of the data path either side of the generator,
through the C ABI: wav payload -> waveform (decode, mono, sox norm, resample), waveform -> PCM16
wav, the fp64 running model average, and the stage-1 pre-training iteration built on them.
The bodies live in tests/_datapath_cases.py (they also run on the CPU against the host-emulated
kernels); named test_zz_* so that this newest file runs after the hot-path suites under `-x`."""
import os

import numpy as np
import pytest
import torch

import _datapath_cases as DC
from _cases import GOLDEN, rel_rms
from oracle import datapath_oracle as DO

pytestmark = pytest.mark.gpu
AVG_TAGS = ["running_fp32", "ema_fp32", "interval_fp64", "running_fp32_acc32", "interval_fp64_acc32"]


def test_load_wav_reproduces_reference_mel_fixture():
    """wav bytes -> f2g_pcm_decode -> LogMel == the reference's fixture mel (test_from_wav.py path)."""
    from flow2gan_b200.modules import LogMelSpectrogram
    audio, g = DC.case_load_fixture("cuda")
    mel = LogMelSpectrogram(24000, 1024, 256, 100).cuda()(audio[None]).cpu()
    assert rel_rms(mel, g["mel"]) < 1e-5


def test_pcm_decode_formats_vs_oracle():
    DC.case_decode_formats("cuda")


@pytest.mark.parametrize("case", DC.G["resample"], ids=lambda c: f"{c['orig']}-{c['new']}-{c['x'].numel()}")
def test_resample_matches_torchaudio_golden(case):
    DC.case_resample_golden("cuda", case)


def test_norm_gain_resample_and_collate():
    DC.case_norm_resample_collate("cuda")


def test_recording_dataset_policy(tmp_path):
    DC.case_recording_dataset("cuda", tmp_path)


def test_pcm16_encode_and_save_wav_round_trip(tmp_path):
    DC.case_encode_save_round_trip("cuda", tmp_path)


@pytest.mark.parametrize("tag", AVG_TAGS)
def test_average_state_dict_bit_exact_vs_reference(tag):
    DC.case_average_state_dict("cuda", tag)
    from flow2gan_b200.averaging import average_state_dict
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        average_state_dict({"w": torch.zeros(3, dtype=torch.float64)}, {"w": torch.zeros(3)}, 0.5, 0.5)


def test_model_average_helpers(tmp_path):
    DC.case_model_average_helpers("cuda", tmp_path)
    DC.case_average_checkpoints("cuda", tmp_path)


def test_large_waveform_streams():
    """Full-size property check (10 min of 48 kHz stereo): decode -> resample -> encode keeps the
    length contract and a low-frequency tone's amplitude; peak/RMS statistics match torch reductions."""
    from flow2gan_b200 import _lib as L
    from flow2gan_b200.datapath import encode_pcm16, gain_resample
    n = 48000 * 600
    k = (torch.arange(n, device="cuda", dtype=torch.int64) * 440) % 48000      # exact phase index
    tone = 0.5 * torch.sin(2 * torch.pi * k.float() / 48000.0)
    pcm = (tone * 32767).round().short()
    payload = torch.stack([pcm, pcm], 1).contiguous().view(torch.uint8).reshape(-1)
    mono = torch.empty(n, device="cuda")
    stats = torch.zeros(2, device="cuda")
    L.pcm_decode(payload, 16, 2, 0, n, mono, stats)
    assert torch.equal(mono, pcm.float() / 32768.0)
    assert abs(float(stats[0]) / float(mono.double().pow(2).sum()) - 1) < 1e-3
    assert float(stats[1]) == float(mono.abs().max())
    y = gain_resample(mono, 48000, 24000, stats, -3.0)
    assert y.numel() == n // 2
    assert abs(float(y[1000:-1000].abs().max()) - 10 ** (-3 / 20)) < 2e-3
    q = encode_pcm16(y)
    assert q.dtype == torch.int16 and int(q.abs().max()) <= 32767


def test_pretrain_step_and_model_average():
    """pretrain.py step body: losses finite, parameters move, both lr groups are scheduled,
    model_avg (fp64 deep copy) follows update_averaged_model's formula bit for bit."""
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.pretrainer import FMTrainer
    from _synth import synth_state_dict
    g = torch.load(os.path.join(GOLDEN, "ref_fm_loss_24k.pt"), weights_only=False)
    m = MelAudioGenerator(**get_generator_config(g["model_name"]))
    m.load_state_dict(synth_state_dict(g["sd_spec"], g["sd_seed"]), strict=False)
    m = m.cuda()
    m.estimators[1].lr_scale = 0.5
    tr = FMTrainer(m, average_period=2, rank=0)
    assert sorted(tr.scheduler.base_lrs) == [0.0175, 0.035]
    assert all(v.dtype == torch.float64 for v in tr.model_avg.state_dict().values() if v.is_floating_point())
    audio, lens = g["audio"].cuda(), g["lens"].cuda()
    torch.manual_seed(0)
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    avg0 = {k: v.detach().cpu().clone() for k, v in tr.model_avg.state_dict().items()}
    losses = [float(tr.step(audio, lens)["loss"]) for _ in range(2)]
    assert all(np.isfinite(losses)), losses
    moved = [k for k, v in m.state_dict().items() if not torch.equal(v, before[k])]
    assert len(moved) > 400, len(moved)
    # after batch 2 (average_period 2): avg = avg0 * 0 + cur * 1 through the reference's op sequence
    want = DO.average_state_dict(avg0, {k: v.detach().cpu() for k, v in m.state_dict().items()}, 1 - 2 / 2, 2 / 2)
    for k, v in tr.model_avg.state_dict().items():
        assert torch.equal(v.cpu(), want[k]), k
    for _ in range(6):
        losses.append(float(tr.step(audio, lens)["loss"]))
    assert np.isfinite(losses).all()
    assert tr.batch_idx_train == 8 and tr.scheduler.batch == 8
