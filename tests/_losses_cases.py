"""Test bodies of the fused multi-tensor loss reductions (csrc/losses.cu) shared by
tests/test_zz_losses_gpu.py (device "cuda", product library) and tests/test_losses_cpu.py (device
"cpu", `_lib.loss_terms` redirected to the host-emulated build of the same kernel source).
Expected values: the reference's own expressions (flow2gan/models/gan.py:57-99) evaluated by torch
autograd on the same tensors."""
import torch
import torch.nn.functional as F


def _views(dev, seed):
    """(ref, x) pairs shaped like the discriminators' feature maps: permuted period-major tensors,
    channel-padded slices, halves of a real|fake batch, plus ragged sizes around the 2048-element
    launch chunk."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    pairs = []
    for (b, p, h, c) in ((2, 3, 37, 32), (2, 5, 11, 128), (3, 2, 7, 4)):
        base_r, base_x = rn(b * p, 1, h, c), rn(b * p, 1, h, c)
        as_bhpc = lambda v: v.unflatten(0, (b, p)).squeeze(2).permute(0, 2, 1, 3)       # noqa: E731
        pairs.append((as_bhpc(base_r), as_bhpc(base_x)))
    pad_r, pad_x = rn(2, 9, 13, 4), rn(2, 9, 13, 4)
    pairs.append((pad_r[..., :1], pad_x[..., :1]))                    # conv_post: Co = 1 of a padded 4
    cat = rn(4, 6, 50, 32)
    pairs.append((cat[:2], cat[2:]))                                  # real | fake halves of one batch
    band = rn(2, 10, 40, 32)
    pairs.append((band[:, :, 3:17, :], rn(2, 10, 14, 32)))            # W-band slice vs contiguous
    for n in (1, 2047, 2048, 2049, 5000):
        pairs.append((rn(n), rn(n)))
    pairs.append((rn(3, 700), rn(3, 700)))
    while len(pairs) < 30:                                            # > 24 terms: two launches
        k = len(pairs)
        pairs.append((rn(2, k + 1, 3), rn(2, k + 1, 3)))
    x0 = pairs[0][1]
    pairs.append((x0.clone(), x0))                                    # identical tensors: sign(0) = 0
    return pairs


def case_l1_terms(dev):
    from flow2gan_b200.losses import l1_terms
    pairs = _views(dev, 1)
    refs = [r for r, _ in pairs]
    xs = [x.clone().requires_grad_(True) for _, x in pairs]
    xs2 = [x.detach().clone().requires_grad_(True) for x in xs]
    got = l1_terms(refs, xs)        # refs keep every stride pattern; dense permuted xs keep theirs too
    want = sum(F.l1_loss(r.detach(), x) for r, x in zip(refs, xs2))
    assert abs(float(got) - float(want)) <= 5e-6 * abs(float(want)), (float(got), float(want))
    (got * 0.37).backward()
    (want * 0.37).backward()
    for i, (a, b) in enumerate(zip(xs, xs2)):
        assert a.grad.shape == b.grad.shape
        assert torch.allclose(a.grad, b.grad, rtol=2e-6, atol=0), i
    assert float(xs[-1].grad.abs().max()) == 0.0


def case_l1_terms_strided_grad_flow(dev):
    """The x_i are non-contiguous VIEWS of leaves (as in the discriminators): autograd must route the
    contiguous gradients back through the views."""
    from flow2gan_b200.losses import l1_terms
    g = torch.Generator().manual_seed(3)
    leaf = torch.randn(6, 1, 20, 8, generator=g).to(dev).requires_grad_(True)
    leaf2 = leaf.detach().clone().requires_grad_(True)
    ref = torch.randn(2, 20, 3, 5, generator=g).to(dev)

    def view(v):
        return v.unflatten(0, (2, 3)).squeeze(2).permute(0, 2, 1, 3)[..., :5]
    l1_terms([ref], [view(leaf)]).backward()
    F.l1_loss(ref, view(leaf2)).backward()
    assert torch.allclose(leaf.grad, leaf2.grad, rtol=2e-6, atol=0)
    assert float(leaf.grad[..., 5:].abs().max()) == 0.0


def case_hinge_terms(dev):
    from flow2gan_b200.losses import hinge_terms
    g = torch.Generator().manual_seed(2)
    shapes = [(4, 298), (4, 1, 47, 130), (2, 3000), (1,), (2, 149, 2, 1)] * 6            # 30 terms
    scores = [(torch.randn(*s, generator=g) * 1.5).to(dev) for s in shapes]
    scores[3] = torch.tensor([-1.0]).to(dev)                          # 1 + s == 0: clamp passes the gradient
    scores[4] = scores[4][..., :1].permute(0, 2, 1, 3)                # a strided score view
    signs = [(-1.0 if i % 2 == 0 else 1.0) for i in range(len(scores))]
    signs[3] = 1.0
    a = [s.clone().requires_grad_(True) for s in scores]
    b = [s.clone().requires_grad_(True) for s in scores]
    got = hinge_terms(a, signs)
    want = sum(torch.mean(torch.clamp(1 + sg * s, min=0)) for s, sg in zip(b, signs))
    assert abs(float(got) - float(want)) <= 5e-6 * abs(float(want))
    (got * 1.7).backward()
    (want * 1.7).backward()
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.allclose(x.grad, y.grad, rtol=2e-6, atol=0), i
    assert float(a[3].grad) != 0.0


def case_gan_loss_methods(dev, monkeypatch):
    """GAN.discriminator_loss / generator_loss / feature_matching_loss with the fused switch on
    against the same static methods with it off (the reference's expressions)."""
    import flow2gan_b200.gan as G
    g = torch.Generator().manual_seed(5)
    sr = [torch.randn(3, 100 + i, generator=g).to(dev) for i in range(8)]
    sf = [torch.randn(3, 100 + i, generator=g).to(dev) for i in range(8)]
    fr = [[torch.randn(3, 4 + j, 5, 8, generator=g).to(dev) for j in range(5)] for _ in range(8)]
    ff = [[torch.randn(3, 4 + j, 5, 8, generator=g).to(dev) for j in range(5)] for _ in range(8)]

    def run(fused):
        monkeypatch.setattr(G, "FUSED_LOSSES", fused)
        a = [t.clone().requires_grad_(True) for t in sr]
        b = [t.clone().requires_grad_(True) for t in sf]
        c = [[t.clone().requires_grad_(True) for t in row] for row in ff]
        d = G.GAN.discriminator_loss(a, b)
        gl = G.GAN.generator_loss(b)
        fm = G.GAN.feature_matching_loss(fr, c)
        (d + 0.5 * gl + 2.0 * fm).backward()
        return (float(d), float(gl), float(fm)), [t.grad for t in a + b] + [t.grad for row in c for t in row]

    v0, g0 = run(False)
    v1, g1 = run(True)
    for x, y in zip(v0, v1):
        assert abs(x - y) <= 3e-6 * abs(x), (v0, v1)
    for i, (x, y) in enumerate(zip(g0, g1)):
        assert torch.allclose(x, y, rtol=3e-6, atol=1e-12), i
