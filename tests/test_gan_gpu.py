"""GAN stage parity against golden outputs of the reference itself (tests/golden/ref_gan_24k.pt):
D-phase / G-phase loss tuples (gan.py:101-166) and the gradients of the stepped half after the
finetune.py loss weighting, plus conv2d building-block checks against torch on CPU.
Loss values: 2e-3 relative (TF32 discriminators + generator); gradients: see test_train_gpu.py."""
import os
import random

import pytest
import torch

from _cases import GOLDEN, rel_rms
from _synth import synth_state_dict
from test_train_gpu import _assert_grads, _grad_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,k,s,p,leaky", [
    ((2, 40, 3, 32), (5, 1), (3, 1), (2, 0), 0.1),
    ((2, 9, 37, 2), (3, 9), (1, 1), (1, 4), 0.1),
    ((2, 9, 37, 32), (3, 9), (1, 2), (1, 4), 0.1),
    ((2, 7, 11, 32), (3, 3), (1, 1), (1, 1), None),
    ((3, 50, 2, 1), (5, 1), (3, 1), (2, 0), 0.1),
    ((2, 12, 5, 1024), (3, 1), (1, 1), (1, 0), None),
])
def test_conv2d_cl_forward_backward(shape, k, s, p, leaky):
    from flow2gan_b200.discriminators import conv2d_cl
    gen = torch.Generator().manual_seed(sum(shape))
    Nb, H, W, C = shape
    Co = 1 if leaky is None else 48
    torch.manual_seed(1000 + sum(shape))                       # deterministic conv initialisation
    conv = torch.nn.Conv2d(C, Co, k, s, padding=p)
    x = torch.randn(Nb, H, W + 6, C, generator=gen)            # a band [3, 3+W) of a wider tensor
    xg = x.cuda().requires_grad_(True)
    conv_g = torch.nn.Conv2d(C, Co, k, s, padding=p).cuda()
    conv_g.load_state_dict(conv.state_dict())
    y = conv2d_cl(xg[:, :, 3:3 + W, :], conv_g, leaky)
    xr = x.clone().requires_grad_(True)
    yr = conv(xr[:, :, 3:3 + W, :].permute(0, 3, 1, 2))
    if leaky is not None:
        yr = torch.nn.functional.leaky_relu(yr, leaky)
    yr = yr.permute(0, 2, 3, 1)
    assert y.shape == yr.shape
    e_fwd = rel_rms(y.detach().cpu(), yr.detach())
    w = torch.randn(yr.shape, generator=gen)
    (y * w.cuda()).sum().backward()
    (yr * w).sum().backward()
    e_dx = rel_rms(xg.grad.cpu(), xr.grad)
    e_dw = rel_rms(conv_g.weight.grad.cpu(), conv.weight.grad)
    e_db = rel_rms(conv_g.bias.grad.cpu(), conv.bias.grad)
    print("conv2d_cl", shape, k, s, "fwd %.2e dx %.2e dw %.2e db %.2e" % (e_fwd, e_dx, e_dw, e_db))
    # backward: a TF32-level forward perturbation flips the LeakyReLU branch of the ~|z|<1e-3 elements,
    # which changes their gate by 10x -> O(sqrt(eps)) ~ 1e-2 gradient differences on low-fan-in convs
    assert e_fwd < 1e-3 and e_dx < 3e-2 and e_dw < 3e-2 and e_db < 3e-2


@pytest.mark.parametrize("Nb,H,C,Co,k,s,leaky", [
    (6, 200, 32, 128, 5, 3, 0.1),          # DiscriminatorP conv 1 (period-major layout)
    (4, 67, 128, 512, 5, 3, 0.1),
    (3, 23, 512, 1024, 5, 3, 0.1),
    (3, 9, 1024, 1024, 5, 1, 0.1),
    (3, 9, 1024, 1, 3, 1, None),           # conv_post
    (5, 300, 1, 32, 5, 3, 0.1),            # conv 0 (C = 1): gather path with swapped axes
])
def test_conv2d_windowed_period_major(Nb, H, C, Co, k, s, leaky):
    """(k, 1) stride (s, 1) Conv2d applied along the contiguous axis of a (Nb, 1, H, C) tensor
    (flow2gan/models/discriminators.py:65-76 on the period-major layout) vs torch fp32 on CPU."""
    from flow2gan_b200 import convwin
    from flow2gan_b200.discriminators import conv2d_cl
    gen = torch.Generator().manual_seed(Nb * 1000 + H + C)
    torch.manual_seed(77 + H + C)
    conv = torch.nn.Conv2d(C, Co, (k, 1), (s, 1), padding=(k // 2, 0))
    x = torch.randn(Nb, 1, H, C, generator=gen)
    xg = x.cuda().requires_grad_(True)
    conv_g = torch.nn.Conv2d(C, Co, (k, 1), (s, 1), padding=(k // 2, 0)).cuda()
    conv_g.load_state_dict(conv.state_dict())
    assert convwin.supports(C, 1, k, 1, s) == (C % 32 == 0)
    y = conv2d_cl(xg, conv_g, leaky, swap_hw=True)                       # (Nb, 1, Ho, Co)
    xr = x.clone().requires_grad_(True)
    yr = conv(xr.permute(0, 3, 2, 1))                                    # (Nb, Co, Ho, 1)
    if leaky is not None:
        yr = torch.nn.functional.leaky_relu(yr, leaky)
    yr = yr.permute(0, 3, 2, 1)
    assert y.shape == yr.shape, (y.shape, yr.shape)
    e_fwd = rel_rms(y.detach().cpu(), yr.detach())
    w = torch.randn(yr.shape, generator=gen)
    (y * w.cuda()).sum().backward()
    (yr * w).sum().backward()
    e_dx = rel_rms(xg.grad.cpu(), xr.grad)
    e_dw = rel_rms(conv_g.weight.grad.cpu(), conv.weight.grad)
    e_db = rel_rms(conv_g.bias.grad.cpu(), conv.bias.grad)
    print("conv2d_win", (Nb, H, C, Co, k, s), "fwd %.2e dx %.2e dw %.2e db %.2e" % (e_fwd, e_dx, e_dw, e_db))
    assert e_fwd < 1e-3 and e_dx < 3e-2 and e_dw < 3e-2 and e_db < 3e-2


@pytest.mark.parametrize("shape,k,s,p,Co", [
    ((4, 47, 103, 32), (3, 9), (1, 2), (1, 4), 32),      # DiscriminatorR convs 1-3, several M tiles
    ((2, 20, 52, 32), (3, 9), (1, 2), (1, 4), 32),       # even width (left-over column)
    ((4, 30, 26, 32), (3, 3), (1, 1), (1, 1), 32),
    ((2, 5, 4, 64), (3, 9), (1, 2), (1, 4), 96),         # kernel wider than the input
])
def test_conv2d_windowed_vs_gather_path(shape, k, s, p, Co):
    """The windowed (implicit im2col) and the gather (materialised im2col) paths of conv2d_cl are two
    implementations of the same Conv2d: outputs and all three gradients must agree to TF32 rounding."""
    from flow2gan_b200 import discriminators as D
    gen = torch.Generator().manual_seed(sum(shape))
    Nb, H, W, C = shape
    torch.manual_seed(5 + sum(shape))
    conv = torch.nn.Conv2d(C, Co, k, s, padding=p).cuda()
    x = torch.randn(Nb, H, W, C, generator=gen).cuda()
    res = []
    for windowed in (True, False):
        D._USE_WINDOWED = windowed
        try:
            xg = x.clone().requires_grad_(True)
            conv.zero_grad()
            y = D.conv2d_cl(xg, conv, 0.1)
            if not res:
                w = torch.randn(y.shape, generator=gen).cuda()
            (y * w).sum().backward()
            res.append((y.detach().clone(), xg.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()))
        finally:
            D._USE_WINDOWED = True
    names = ("fwd", "dx", "dw", "db")
    errs = {n: rel_rms(a, b) for n, a, b in zip(names, res[0], res[1])}
    print("windowed vs gather", shape, errs)
    # identical TF32 products, different summation order; the forward rounding differences flip a few
    # LeakyReLU branches in the backward (see above)
    assert errs["fwd"] < 1e-5 and errs["dx"] < 1e-2 and errs["dw"] < 1e-2 and errs["db"] < 1e-2


def _gan(g):
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    gen = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    gen.branch_dropout = 0.0                                   # finetune.py:414
    gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    missing, unexpected = gan.load_state_dict(synth_state_dict(g["sd_spec"], g["sd_seed"]), strict=False)
    assert not unexpected
    return gan.cuda()


@pytest.mark.parametrize("tag,draw", [("limit_on", 0.0), ("limit_off", 0.99)])
def test_gan_phases_match_reference(tag, draw):
    g = torch.load(os.path.join(GOLDEN, "ref_gan_24k.pt"), weights_only=False)
    gan = _gan(g)
    from flow2gan_b200.modules import LogMelSpectrogram
    audio, lens = g["audio"].cuda(), g["lens"].cuda()
    mel = LogMelSpectrogram(24000, 1024, 256, 100).cuda()(audio)
    assert rel_rms(mel.cpu(), g["mel"]) < 1e-5
    noise = g["noise"].cuda()
    rr = random.random
    random.random = lambda: draw                                # pins limit_param_value (modules.py:267)
    try:
        for train_disc, ph, wts in ((True, "d", (1.0, 0.1)), (False, "g", (1.0, 0.1, 1.0, 0.1, 45.0))):
            gan.zero_grad()
            losses = gan(cond=mel, audio=audio, audio_lens=lens, n_timesteps=1, train_disc=train_disc,
                         noise=noise)
            got = torch.stack([l.detach() for l in losses]).cpu()
            ref = g[f"{ph}_{tag}_losses"]
            rel = ((got - ref).abs() / ref.abs()).max()
            print(ph, tag, "losses", got.tolist(), "max rel", float(rel))
            assert float(rel) < 2e-3
            assert gan.discriminator.training == train_disc and gan.generator.training == (not train_disc)
            sum(l * w for l, w in zip(losses, wts)).backward()
            sub = gan.discriminator if train_disc else gan.generator
            pre = "discriminator." if train_disc else "generator."
            errs = _grad_errors([(pre + k, p) for k, p in sub.named_parameters()], g[f"{ph}_{tag}_grads"])
            _assert_grads(errs)
    finally:
        random.random = rr
