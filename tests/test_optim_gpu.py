"""Fused multi-tensor ScaledAdam (csrc/optim.cu) + Eden2 vs. the reference's own optimizer run
(tests/golden/ref_scaled_adam.pt) and vs. the CPU oracle on a second random problem."""
import os

import pytest
import torch

from _cases import GOLDEN, rel_rms
from oracle import flow2gan_oracle as O

pytestmark = pytest.mark.gpu


def _run(g, upto=None, resume_at=None):
    from flow2gan_b200.optim import Eden2, ScaledAdam
    h = g["hyper"]
    names = [n for n, _ in g["shapes"]]
    params = [torch.nn.Parameter(p.clone().cuda()) for p in g["init"]]
    opt = ScaledAdam(list(zip(names, params)), lr=h["lr"], clipping_scale=h["clipping_scale"])
    sched = Eden2(opt, lr_batches=h["lr_batches"], warmup_batches=h["warmup_batches"],
                  warmup_start=h["warmup_start"])
    lrs = []
    for step, gs in enumerate(g["grads"][:upto]):
        if resume_at is not None and step == resume_at:
            sd, ssd = opt.state_dict(), sched.state_dict()
            params = [torch.nn.Parameter(p.detach().clone()) for p in params]
            opt = ScaledAdam(list(zip(names, params)), lr=h["lr"], clipping_scale=h["clipping_scale"])
            sched = Eden2(opt, lr_batches=h["lr_batches"], warmup_batches=h["warmup_batches"],
                          warmup_start=h["warmup_start"])
            opt.load_state_dict(sd)
            sched.load_state_dict(ssd)
            sched._set_lrs()
        for p, gr in zip(params, gs):
            p.grad = None if gr is None else gr.clone().cuda()
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step_batch()
    return params, lrs, opt


def test_scaled_adam_matches_reference_run():
    g = torch.load(os.path.join(GOLDEN, "ref_scaled_adam.pt"), weights_only=False)
    params, lrs, opt = _run(g)
    assert max(abs(a - b) for a, b in zip(lrs, g["lrs"])) < 1e-12
    for p, ref, (n, _) in zip(params, g["final"], g["shapes"]):
        assert rel_rms(p.detach().cpu(), ref) < 2e-5, n
    sd = opt.state_dict()
    st = sd["state"]
    assert any("model_norms" in v for v in st.values())
    assert all(set(v) >= {"step", "exp_avg_sq", "delta"} for v in st.values())


def test_scaled_adam_checkpoint_resume_is_transparent():
    g = torch.load(os.path.join(GOLDEN, "ref_scaled_adam.pt"), weights_only=False)
    a, _, _ = _run(g, upto=30)
    b, _, _ = _run(g, upto=30, resume_at=17)
    for x, y in zip(a, b):
        assert rel_rms(x.detach().cpu(), y.detach().cpu()) < 1e-6


def test_scaled_adam_many_tensors_vs_oracle():
    from flow2gan_b200.optim import ScaledAdam
    gen = torch.Generator().manual_seed(11)
    shapes = [(64, 33, 3)] * 5 + [(257,)] * 7 + [()] * 9 + [(300, 1)] * 3 + [(5000, 9)]
    names = [f"p{i}" for i in range(len(shapes))]
    init = [torch.randn(s, generator=gen) * 0.2 for s in shapes]
    cpu = [p.clone() for p in init]
    gpu = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    ora = O.ScaledAdamOracle(names, cpu, lr=0.01, clipping_scale=2.0)
    opt = ScaledAdam(list(zip(names, gpu)), lr=0.01, clipping_scale=2.0)
    for step in range(24):
        gs = [torch.randn(s, generator=gen) * (8.0 if step == 15 else 1.0) for s in shapes]
        ora.step(cpu, gs)
        for p, gr in zip(gpu, gs):
            p.grad = gr.cuda()
        opt.step()
    for a, b, n in zip(gpu, cpu, names):
        assert rel_rms(a.detach().cpu(), b) < 2e-5, n
