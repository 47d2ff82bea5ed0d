"""Direct first-layer convolutions of the discriminators (csrc/conv.cu::conv_small_*; Cin = 1 / 2 -> 32
channels + LeakyReLU, flow2gan/models/discriminators.py:65,171) on the host emulation of the same source,
through the product's autograd function, against torch.nn.functional.conv2d: forward and all three
gradients, strided band views, the swapped-axes (k,1) form of DiscriminatorP, and the routing."""
import pytest
import torch

import _emul
from _cases import rel_rms

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


@pytest.fixture
def L(monkeypatch):
    return _emul.native_fixture(monkeypatch)


CASES = [
    # (Nb, H, W, Cin), (kh, kw), (sh, sw), (ph, pw), leaky, band offset in a wider tensor
    ((2, 7, 23, 2), (3, 9), (1, 1), (1, 4), 0.1, 3),        # DiscriminatorR conv 0 on a frequency band
    ((3, 5, 9, 2), (3, 9), (1, 1), (1, 4), 0.1, 0),         # kernel as wide as the band
    ((4, 1, 50, 1), (1, 5), (1, 3), (0, 2), 0.1, 0),        # DiscriminatorP conv 0, period-major (swapped axes)
    ((2, 1, 17, 1), (1, 5), (1, 3), (0, 2), None, 0),       # no activation
    ((2, 6, 8, 1), (3, 3), (2, 1), (1, 1), 0.2, 0),         # strided rows
]


@pytest.mark.parametrize("shape,k,s,p,leaky,off", CASES)
def test_conv_small_matches_torch_conv2d(L, shape, k, s, p, leaky, off):
    from flow2gan_b200.discriminators import _ConvSmallFn, _small_ok
    Nb, H, W, Cin = shape
    g = torch.Generator().manual_seed(sum(shape) + k[1])
    wide = torch.randn(Nb, H, W + 2 * off, Cin, generator=g)
    weight = torch.randn(32, Cin, *k, generator=g) * 0.3
    bias = torch.randn(32, generator=g) * 0.1
    xa = wide.clone().requires_grad_(True)
    wa, ba = weight.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    xv = xa[:, :, off:off + W, :]
    assert _small_ok(xv, wa)
    y = _ConvSmallFn.apply(xv, wa, ba, s[0], s[1], p[0], p[1], leaky)
    xr = wide.clone().requires_grad_(True)
    wr, br = weight.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr[:, :, off:off + W, :].permute(0, 3, 1, 2), wr, br, stride=s, padding=p)
    if leaky is not None:
        yr = torch.nn.functional.leaky_relu(yr, leaky)
    yr = yr.permute(0, 2, 3, 1)
    assert y.shape == yr.shape
    assert rel_rms(y.detach(), yr.detach()) < 1e-6
    gy = torch.randn(yr.shape, generator=g)
    (y * gy).sum().backward()
    (yr * gy).sum().backward()
    assert rel_rms(xa.grad, xr.grad) < 1e-5
    assert rel_rms(wa.grad, wr.grad) < 1e-5
    assert rel_rms(ba.grad, br.grad) < 1e-5


def test_first_layers_take_the_direct_path(L, monkeypatch):
    """conv2d_cl routes Cin = 1 / 2 -> 32 convs to the direct kernels (no im2col launch), also with frozen
    weights (G phase: gradient w.r.t. the input only) and swapped axes."""
    import flow2gan_b200.discriminators as D
    calls = []
    monkeypatch.setattr(L, "im2col2d", lambda *a, **k: calls.append("im2col"))
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(1, 32, (5, 1), (3, 1), padding=(2, 0))
    x = torch.randn(6, 1, 40, 1, requires_grad=True)
    y = D.conv2d_cl(x, conv, 0.1, train_weights=False, swap_hw=True)
    yr = torch.nn.functional.leaky_relu(conv(x.detach().permute(0, 3, 2, 1)), 0.1).permute(0, 3, 2, 1)
    assert not calls and rel_rms(y.detach(), yr.detach()) < 1e-6
    y.square().sum().backward()
    assert x.grad is not None and conv.weight.grad is None
