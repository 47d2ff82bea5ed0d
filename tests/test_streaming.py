"""Chunked ("streaming") synthesis, flow2gan/bin/infer_dir.py:126-168: chunk geometry on CPU;
on the GPU the chunked output against the oracle run chunk by chunk with the same pinned noise,
and the batched-chunk extension against the sequential mode."""
import pytest
import torch

from _cases import mel_input, noise_input, rel_rms
from flow2gan_b200.config import AttributeDict
from flow2gan_b200.streaming import SIDE_CONTEXT_FRAMES, chunk_plan


@pytest.mark.parametrize("frames,chunk", [(150, 50), (205, 64), (40, 100), (100, 100), (101, 100), (1203, 200)])
def test_chunk_plan_geometry(frames, chunk):
    hop = 256
    plan = chunk_plan(frames, chunk, hop)
    assert len(plan) == (frames + chunk - 1) // chunk
    total = 0
    for i, (f0, f1, lp, rp) in enumerate(plan):
        assert f0 == max(0, i * chunk - SIDE_CONTEXT_FRAMES) and f1 == min(frames, (i + 1) * chunk + SIDE_CONTEXT_FRAMES)
        size = (f1 - f0) * hop
        kept = len(range(size)[lp: size - rp])           # the reference's slice, negative rp included
        assert kept == (min(frames, (i + 1) * chunk) - i * chunk) * hop
        total += kept
    assert total == frames * hop                          # chunks tile the utterance exactly
    assert len({f1 - f0 for f0, f1, _, _ in plan}) <= 4   # few launch-graph shapes (plan cache holds 8)


@pytest.mark.gpu
def test_streaming_matches_oracle_chunk_by_chunk():
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.streaming import infer_audio, streaming_infer_audio
    from oracle import flow2gan_oracle as O
    from _synth import synth_state_dict
    torch.manual_seed(0)
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    B, frames, chunk, hop = 2, 150, 50, 256
    mel = mel_input(B, 100, frames, seed=5)
    params = AttributeDict(n_timesteps=2, chunk_size=chunk)
    plan = chunk_plan(frames, chunk, hop)
    noises = [noise_input(B, (f1 - f0) * hop, seed=40 + i) for i, (f0, f1, _, _) in enumerate(plan)]
    got = streaming_infer_audio(params, m, None, cond=mel, noise_fn=lambda i, shape: noises[i])
    assert got.shape == (B, frames * hop) and got.device.type == "cpu"
    cfg = O.generator_config("mel_24k_base")
    ref = []
    with torch.no_grad():
        for (f0, f1, lp, rp), nz in zip(plan, noises):
            a = O.generator_infer(sd, cfg, mel[:, :, f0:f1], nz, None, 2, True)
            ref.append(a[:, lp: a.size(1) - rp])
    ref = torch.cat(ref, -1)
    err = rel_rms(got, ref)
    print("streaming rel-RMS vs oracle:", err)
    assert err < 1e-3
    # interior chunks stacked along the batch axis: same numbers, fewer calls
    got_b = streaming_infer_audio(params, m, None, cond=mel, batch_chunks=True,
                                  noise_fn=lambda i, shape: noises[i])
    assert rel_rms(got_b, got) < 1e-5
    # default noise path: same global-RNG order as a hand-written loop over model.infer
    torch.manual_seed(7)
    a = streaming_infer_audio(params, m, None, cond=mel)
    torch.manual_seed(7)
    with torch.inference_mode():
        parts = []
        for f0, f1, lp, rp in plan:
            p = m.infer(cond=mel[:, :, f0:f1].cuda().contiguous(), n_timesteps=2, clamp_pred=True)
            parts.append(p[:, lp: p.size(1) - rp])
    assert torch.equal(a, torch.cat(parts, -1).cpu())
    whole = infer_audio(params, m, None, cond=mel)
    assert whole.shape == got.shape and float(whole.abs().max()) <= 1.0
