"""Parity at BASELINE.json's TRAIN configuration (configs[2]: mel_24k_base GAN fine-tune step, bs 16 x
24000 samples): both phases' loss tuples and the gradients of the stepped half against autograd of the
fp32 CPU oracle on the same seeded batch.  (tests/test_gan_gpu.py checks the same quantities against
golden outputs of the reference itself at bs 2 x 6144; the oracle is pinned to the reference by
tests/test_oracle_vs_golden.py.)  The oracle passes take ~1 min on the GPU box's host cores."""
import random

import pytest
import torch

from _cases import audio_input, noise_input, rel_rms
from _synth import synth_state_dict
from oracle import flow2gan_oracle as O
from test_train_gpu import _assert_grads

pytestmark = pytest.mark.gpu
B, T = 16, 24000
W_D, W_G = (1.0, 0.1), (1.0, 0.1, 1.0, 0.1, 45.0)              # finetune.py loss weights


def _errs(named_params, ref_grads):
    out = {}
    for k, p in named_params:
        r = ref_grads.get(k)
        if r is None or p.grad is None:
            continue
        l2 = float(r.double().norm())
        if l2 < 1e-6:
            continue
        out[k] = (float((p.grad.detach().float().cpu() - r).double().norm()) / l2, l2)
    return out


def test_gan_phases_at_train_config_vs_oracle():
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.modules import LogMelSpectrogram
    torch.manual_seed(0)
    gen = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    gen.branch_dropout = 0.0                                   # finetune.py:414
    gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in gan.state_dict().items()], 4321)
    gan.load_state_dict(sd, strict=False)
    gan = gan.cuda()
    audio = audio_input(B, T, seed=2)
    lens = torch.full((B,), T, dtype=torch.int64)
    noise = noise_input(B, T, seed=9)
    mel = LogMelSpectrogram(24000, 1024, 256, 100).cuda()(audio.cuda())
    mel_ref = O.log_mel(audio)
    assert rel_rms(mel.cpu(), mel_ref) < 1e-5
    cfg = O.generator_config("mel_24k_base")
    torch.set_num_threads(min(32, torch.get_num_threads() or 1))
    rr = random.random
    random.random = lambda: 0.99                                # limit_param_value off on both sides (modules.py:267)
    try:
        for train_disc, wts, pre in ((True, W_D, "discriminator."), (False, W_G, "generator.")):
            # ---- CUDA path
            gan.zero_grad()
            losses = gan(cond=mel, audio=audio.cuda(), audio_lens=lens.cuda(), n_timesteps=1,
                         train_disc=train_disc, noise=noise.cuda())
            got = torch.stack([l.detach() for l in losses]).cpu()
            sum(l * w for l, w in zip(losses, wts)).backward()
            # ---- fp32 CPU oracle with autograd over the stepped half
            sd_o = {k: (v.clone().requires_grad_(True) if k.startswith(pre) else v.clone()) for k, v in sd.items()}
            for k, v in gan.state_dict().items():              # buffers (windows, filterbanks) as the modules built them
                if k not in sd_o:
                    sd_o[k] = v.detach().cpu()
            ref_losses = O.gan_forward(sd_o, cfg, mel_ref, audio, noise, lens, 1, train_disc)
            ref = torch.stack([l.detach() for l in ref_losses])
            sum(l * w for l, w in zip(ref_losses, wts)).backward()
            rel = ((got - ref).abs() / ref.abs()).max()
            print("train config", "D" if train_disc else "G", "losses", got.tolist(), "oracle", ref.tolist(),
                  "max rel %.2e" % float(rel))
            assert float(rel) < 2e-3
            sub = gan.discriminator if train_disc else gan.generator
            ref_grads = {k: v.grad for k, v in sd_o.items() if v.requires_grad and v.grad is not None}
            errs = _errs([(pre + k, p) for k, p in sub.named_parameters()], ref_grads)
            assert len(errs) > 100, len(errs)
            _assert_grads(errs)
    finally:
        random.random = rr
