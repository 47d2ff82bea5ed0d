"""Backward-pass parity: CUDA generator fwd+bwd (flow2gan_b200.train) against golden gradients of
the reference itself (stage-1 FM loss) and against autograd of the CPU oracle; adjoint identities
for the STFT / iSTFT backward kernels.

Gradient tolerance: the forward pass is TF32 (5e-4 rel-RMS on activations, the parity gate); a
gradient through ~40 serial TF32 contractions accumulates that to ~1-2e-2 per tensor, and scalar
parameters whose gradient is a cancelling sum (BiasNorm log_scale) reach ~6e-2.  Even the fp32
reference vs the fp32 oracle differ by up to 1e-2 on such tensors (test_oracle_vs_golden.py).
Gates: see _assert_grads; loss value itself < 2e-3."""
import os

import pytest
import torch

from _cases import GOLDEN, audio_input, mel_input, noise_input, rel_rms
from oracle import flow2gan_oracle as O
from _synth import synth_state_dict

pytestmark = pytest.mark.gpu


def _model(name, spec, seed):
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**get_generator_config(name))
    sd = synth_state_dict(spec, seed)
    m.load_state_dict(sd, strict=False)
    return m.cuda(), sd


def _grad_errors(named_params, golden):
    """per tensor: (relative error, reference l2)."""
    errs = {}
    for k, p in named_params:
        e = golden[k]
        if e is None:
            continue
        assert p.grad is not None, k
        g = p.grad.detach().float().cpu()
        if e["l2"] < 1e-6:      # exactly-cancelling gradient in the reference (e.g. hinge conv_post.bias)
            assert float(g.double().norm()) < 1e-4, (k, float(g.double().norm()))
            continue
        if "full" in e:
            err = float((g - e["full"]).double().norm()) / e["l2"]
        else:
            smp = g.flatten()[::e["stride"]][:2048]
            err = max(abs(float(g.double().norm()) - e["l2"]) / e["l2"],
                      float((smp - e["sample"]).double().norm() / e["sample"].double().norm().clamp_min(1e-12)))
        errs[k] = (err, e["l2"])
    return errs


def _assert_grads(errs, total=3e-2, p90=5e-2, median=2e-2, worst_big=1.5e-1):
    """errs: name -> (rel err, ref l2).  Gates: the norm-weighted error of the whole gradient
    vector, the median / 90th percentile over tensors, and the worst tensor among those that carry
    >= 1 % of the largest tensor norm (tiny cancellation-dominated tensors -- e.g. a time_embed_proj
    whose gradient is a near-zero sum over frames -- amplify TF32 noise arbitrarily)."""
    v = sorted(e for e, _ in errs.values())
    top = sorted(((e, k) for k, (e, _) in errs.items()), reverse=True)[:8]
    l2max = max(l for _, l in errs.values())
    tot = (sum((e * l) ** 2 for e, l in errs.values()) / sum(l ** 2 for _, l in errs.values())) ** 0.5
    big = max(e for e, l in errs.values() if l >= 1e-2 * l2max)
    print("grad rel-err: whole-vector %.2e  median %.2e  p90 %.2e  max %.2e  max(big tensors) %.2e"
          % (tot, v[len(v) // 2], v[int(len(v) * 0.9)], v[-1], big))
    print("   worst tensors:", [(k, round(e, 4)) for e, k in top])
    assert tot < total and big < worst_big
    assert v[int(len(v) * 0.9)] < p90
    assert v[len(v) // 2] < median


def test_fm_loss_and_grads_match_reference():
    g = torch.load(os.path.join(GOLDEN, "ref_fm_loss_24k.pt"), weights_only=False)
    m, sd = _model(g["model_name"], g["sd_spec"], g["sd_seed"])
    m.eval()          # the golden run used eval(): no branch dropout / limit_param_value draws
    loss = m(cond=g["mel"].cuda(), audio=g["audio"].cuda(), audio_lens=g["lens"].cuda(),
             noise=g["noise"].cuda(), t=g["t"].cuda())
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    print("fm loss", float(loss), "ref", float(g["loss"]), "rel", rel)
    assert rel < 2e-3
    loss.backward()
    _assert_grads(_grad_errors(list(m.named_parameters()), g["grads"]))


def test_two_step_sampler_grads_vs_oracle_autograd():
    """Euler unrolling with n_timesteps=2, train mode with pinned limit_param_value draw, gradient
    w.r.t. parameters AND the initial noise (exercises the STFT adjoint path)."""
    import random
    g = torch.load(os.path.join(GOLDEN, "ref_infer_24k.pt"), weights_only=False)
    m, sd = _model(g["model_name"], g["sd_spec"], 777)
    m.train()
    m.branch_dropout = 0.0
    B, Fm = 2, 10
    mel = mel_input(B, 100, Fm, seed=5)
    noise = noise_input(B, Fm * 256, seed=6)
    tgt = audio_input(B, Fm * 256, seed=7)
    rr = random.random
    random.random = lambda: 0.0          # limit_param_value always applied (modules.py:267)
    try:
        nz = noise.cuda().requires_grad_(True)
        out = m.infer(mel.cuda(), n_timesteps=2, noise=nz)
        loss = ((out - tgt.cuda()) ** 2).mean() * 100
        loss.backward()
    finally:
        random.random = rr
    cfg = O.generator_config(g["model_name"])
    leaves = {k: v.clone().requires_grad_(not (k.endswith("window") or k.endswith(".fb"))) for k, v in sd.items()}
    nz_ref = noise.clone().requires_grad_(True)
    ref = O.generator_infer(leaves, cfg, mel, nz_ref, None, 2, False, limit=True)
    assert rel_rms(out.detach().cpu(), ref.detach()) < 1e-3
    (((ref - tgt) ** 2).mean() * 100).backward()
    errs = {}
    for k, p in m.named_parameters():
        r = leaves[k].grad
        errs[k] = (float((p.grad.cpu() - r).double().norm() / r.double().norm().clamp_min(1e-12)),
                   float(r.double().norm()))
    errs["<noise>"] = (float((nz.grad.cpu() - nz_ref.grad).double().norm() / nz_ref.grad.double().norm()),
                       float(nz_ref.grad.double().norm()))
    _assert_grads(errs)


def test_fm_loss_train_mode_branch_dropout_vs_oracle(monkeypatch):
    """Train-mode branch dropout (generator.py:145-162): with the three draws pinned (which branch, whether
    to drop, per batch element) the CUDA forward + backward must match the oracle's autograd run with the
    same per-sample branch weights -- dropped branches contribute neither to the prediction nor to the
    gradients, the two kept branches are rescaled by 3/2."""
    import random
    import flow2gan_b200.train as TR
    g = torch.load(os.path.join(GOLDEN, "ref_fm_loss_24k.pt"), weights_only=False)
    m, sd = _model(g["model_name"], g["sd_spec"], 4242)
    m.train()
    m.branch_dropout = 0.05
    Bn = g["audio"].shape[0]
    idx = torch.tensor([(2 * i + 1) % 3 for i in range(Bn)])                   # dropped branch per element
    drop = torch.tensor([[0.01 if i % 2 == 0 else 0.9] for i in range(Bn)])     # < 0.05 -> element i drops
    calls = []
    real_randint, real_rand = torch.randint, torch.rand

    def fake_randint(lo, hi, size, **kw):
        calls.append(("randint", tuple(size)))
        assert (lo, hi, tuple(size)) == (0, 3, (Bn,))
        return idx.to(kw.get("device", "cpu"))

    def fake_rand(size, **kw):
        calls.append(("rand", tuple(size)))
        assert tuple(size) == (Bn, 1)
        return drop.to(kw.get("device", "cpu"))

    monkeypatch.setattr(TR.torch, "randint", fake_randint)
    monkeypatch.setattr(TR.torch, "rand", fake_rand)
    rr = random.random
    random.random = lambda: 0.99          # limit_param_value off
    try:
        loss = m(cond=g["mel"].cuda(), audio=g["audio"].cuda(), audio_lens=g["lens"].cuda(),
                 noise=g["noise"].cuda(), t=g["t"].cuda())
        loss.backward()
    finally:
        random.random = rr
        monkeypatch.setattr(TR.torch, "randint", real_randint)
        monkeypatch.setattr(TR.torch, "rand", real_rand)
    assert calls == [("randint", (Bn,)), ("rand", (Bn, 1))], calls              # the reference's draw order
    mask = torch.ones(Bn, 3)
    mask[torch.arange(Bn), idx] = 0.0
    weight = torch.where(drop < 0.05, mask * 1.5, torch.ones_like(mask))
    assert bool((weight == 0).any())
    cfg = O.generator_config(g["model_name"])
    leaves = {k: v.clone().requires_grad_(not (k.endswith("window") or k.endswith(".fb"))) for k, v in sd.items()}
    for k, v in m.state_dict().items():
        if k not in leaves:
            leaves[k] = v.detach().cpu()
    ref = O.fm_loss(leaves, cfg, g["mel"], g["audio"], g["lens"], g["noise"], g["t"], branch_weight=weight)
    rel = abs(float(loss.detach()) - float(ref.detach())) / float(ref.detach())
    print("fm loss with branch dropout", float(loss.detach()), "oracle", float(ref.detach()), "rel", rel)
    assert rel < 2e-3
    ref.backward()
    errs = {}
    for k, p in m.named_parameters():
        r = leaves[k].grad
        if r is None or p.grad is None or float(r.double().norm()) < 1e-9:
            continue
        errs[k] = (float((p.grad.cpu() - r).double().norm() / r.double().norm()), float(r.double().norm()))
    _assert_grads(errs)


@pytest.mark.parametrize("n_fft,hop", [(128, 64), (512, 256), (1024, 256), (64, 16)])
def test_stft_adjoint_identity(n_fft, hop):
    """<STFT x, G> == <x, STFT^T G> and the same for the iSTFT pair (size-independent property)."""
    from flow2gan_b200 import _lib as L
    B, T = 2, 3000
    gen = torch.Generator().manual_seed(n_fft + hop)
    x = torch.randn(B, T, generator=gen).cuda()
    F = 1 + T // hop
    ld = n_fft + 4
    G = torch.zeros(B * F, ld)
    G[:, : n_fft + 2] = torch.randn(B * F, n_fft + 2, generator=gen)
    G = G.cuda()
    S = torch.zeros(B * F, ld, device="cuda")
    L.stft(x, B, T, T, n_fft, hop, L.SPEC_PACKED, S, ld)
    fr = torch.empty(B * F, n_fft, device="cuda")
    L.stft_bwd_frames(G, B * F, ld, n_fft, fr)
    dx = torch.empty(B, T, device="cuda")
    L.stft_bwd_fold(fr, B, T, n_fft, hop, F, dx, False)
    lhs = float((S.double() * G.double()).sum())
    rhs = float((x.double() * dx.double()).sum())
    assert abs(lhs - rhs) < 2e-5 * max(abs(lhs), 1.0), (lhs, rhs)
    if hop * 2 == n_fft:
        # iSTFT: y = istft(P);  <y, g> == <P, istft^T g>
        P = G.clone()
        frs = torch.empty(B * F, n_fft, device="cuda")
        L.irfft_frames(P, B * F, ld, n_fft, frs)
        y = torch.empty(B, T, device="cuda")
        L.ola_combine([frs], [n_fft], [hop], [F], None, None, y, B, T, False, 0.0, 0.0, False)
        gy = torch.randn(B, T, generator=gen).cuda()
        Lp = n_fft + hop * (F - 1)
        gs = torch.empty(B, Lp, device="cuda")
        L.istft_bwd_prep(gy, B, T, n_fft, hop, F, 1.0, gs)
        dP = torch.zeros(B * F, ld, device="cuda")
        L.istft_bwd_spec(gs, B, Lp, n_fft, hop, None, dP, ld)
        nb = n_fft // 2 + 1
        Pm = P.clone()
        Pm[:, nb] = 0           # Im(DC), Im(Nyquist) are ignored by the C2R transform
        Pm[:, 2 * nb - 1] = 0
        lhs = float((y.double() * gy.double()).sum())
        rhs = float((Pm.double() * dP.double()).sum())
        assert abs(lhs - rhs) < 2e-5 * max(abs(lhs), 1.0), (lhs, rhs)


def test_filterbank_loss_backward_vs_oracle():
    from flow2gan_b200 import _lib as L
    from flow2gan_b200.losses import filter_spec_rows
    from flow2gan_b200.modules import linear_fbanks, melscale_fbanks
    B, T = 2, 4000
    x = audio_input(B, T, seed=3)
    for (n, hop, mode, fb, clip) in ((1024, 256, L.SPEC_POWER, linear_fbanks(513, 256, 24000), 0.0),
                                     (256, 64, L.SPEC_MAG, melscale_fbanks(129, 40, 24000), 1e-7),
                                     (32, 8, L.SPEC_MAG, melscale_fbanks(17, 5, 24000), 1e-7)):
        xg = x.cuda().requires_grad_(True)
        out = filter_spec_rows(xg, fb.cuda(), n, hop, mode, clip)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(n)).cuda()
        (out * w).sum().backward()
        xr = x.clone().requires_grad_(True)
        s = O.stft_complex(xr, n, hop).abs()
        s = s.pow(2.0) if mode == L.SPEC_POWER else s
        f = torch.matmul(s.transpose(1, 2), fb)                 # (B, F, n_filt)
        f = O.safe_log(f) if clip > 0 else f
        assert rel_rms(out.detach().cpu().view(f.shape), f.detach()) < 1e-4
        (f * w.cpu().view(f.shape)).sum().backward()
        assert rel_rms(xg.grad.cpu(), xr.grad) < 2e-4, (n, hop)
