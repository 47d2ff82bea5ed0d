"""GPU parity of the fused multi-tensor loss reductions (csrc/losses.cu) through the C ABI against
torch autograd of the reference's expressions (flow2gan/models/gan.py:57-99); bodies shared with the
CPU run on the host-emulated kernels (tests/_losses_cases.py).  Named test_zz_* so that this newest
file runs after the hot-path suites under `-x`."""
import pytest
import torch

import _losses_cases as LC

pytestmark = pytest.mark.gpu


def test_l1_terms():
    LC.case_l1_terms("cuda")
    LC.case_l1_terms_strided_grad_flow("cuda")


def test_hinge_terms():
    LC.case_hinge_terms("cuda")


def test_gan_loss_methods(monkeypatch):
    LC.case_gan_loss_methods("cuda", monkeypatch)


def test_fused_losses_inside_gan_forward_match_unfused(monkeypatch):
    """Both GAN phases on the golden GAN case with F2G_FUSED_LOSSES on vs off: same loss tuples, same
    gradients up to the summation order of the reductions (whole-vector rel-RMS; the split-K
    weight-gradient atomics and TF32 rounding of the propagated gradients add ~1e-5 of run-to-run
    noise on their own)."""
    import os
    import random
    import flow2gan_b200.gan as G
    from _cases import GOLDEN
    from _synth import synth_state_dict
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.modules import LogMelSpectrogram
    g = torch.load(os.path.join(GOLDEN, "ref_gan_24k.pt"), weights_only=False)
    gen = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    gen.branch_dropout = 0.0
    gan = G.GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    gan.load_state_dict(synth_state_dict(g["sd_spec"], g["sd_seed"]), strict=False)
    gan = gan.cuda()
    audio, lens, noise = g["audio"].cuda(), g["lens"].cuda(), g["noise"].cuda()
    cond = LogMelSpectrogram(24000, 1024, 256, 100).cuda()(audio)
    results = {}
    rr = random.random
    random.random = lambda: 0.99                      # limit_param_value hook off (modules.py:267)
    try:
        for fused in (False, True):
            monkeypatch.setattr(G, "FUSED_LOSSES", fused)
            out = {}
            for disc, wts in ((True, (1.0, 0.1)), (False, (1.0, 0.1, 1.0, 0.1, 45.0))):
                gan.zero_grad(set_to_none=True)
                losses = gan(cond=cond, audio=audio, audio_lens=lens, n_timesteps=1, train_disc=disc, noise=noise)
                sum(l * w for l, w in zip(losses, wts)).backward()
                half = gan.discriminator if disc else gan.generator
                out[disc] = ([float(l.detach()) for l in losses],
                             {k: p.grad.detach().clone() for k, p in half.named_parameters() if p.grad is not None})
            results[fused] = out
    finally:
        random.random = rr
    for disc in (True, False):
        l0, g0 = results[False][disc]
        l1, g1 = results[True][disc]
        for a, b in zip(l0, l1):
            assert abs(a - b) <= 1e-5 * max(abs(a), 1e-6), (disc, l0, l1)
        assert g0.keys() == g1.keys()
        num = sum(float((g0[k] - g1[k]).double().pow(2).sum()) for k in g0)
        den = sum(float(g0[k].double().pow(2).sum()) for k in g0)
        assert (num / den) ** 0.5 < 1e-3, (disc, (num / den) ** 0.5)
