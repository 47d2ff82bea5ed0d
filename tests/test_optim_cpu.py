"""ScaledAdam on the CPU: the kernels of csrc/optim.cu compiled for the host from the same source
(cooperative emulation: one host thread per CUDA thread, tests/_emul.py) behind the product's
optimizer class, against the reference's own 45-step optimizer run (tests/golden/ref_scaled_adam.pt)
and the CPU oracle.  Same cases as tests/test_optim_gpu.py."""
import os

import pytest
import torch

import _emul
from _cases import GOLDEN, rel_rms
from oracle import flow2gan_oracle as O

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


@pytest.fixture
def emulated_adam(monkeypatch):
    return _emul.native_fixture(monkeypatch)


def _run(g, upto=None, resume_at=None):
    from flow2gan_b200.optim import Eden2, ScaledAdam
    h = g["hyper"]
    names = [n for n, _ in g["shapes"]]
    params = [torch.nn.Parameter(p.clone()) for p in g["init"]]

    def make(ps):
        opt = ScaledAdam(list(zip(names, ps)), lr=h["lr"], clipping_scale=h["clipping_scale"])
        return opt, Eden2(opt, lr_batches=h["lr_batches"], warmup_batches=h["warmup_batches"],
                          warmup_start=h["warmup_start"])
    opt, sched = make(params)
    lrs = []
    for step, gs in enumerate(g["grads"][:upto]):
        if resume_at is not None and step == resume_at:
            sd, ssd = opt.state_dict(), sched.state_dict()
            params = [torch.nn.Parameter(p.detach().clone()) for p in params]
            opt, sched = make(params)
            opt.load_state_dict(sd)
            sched.load_state_dict(ssd)
            sched._set_lrs()
        for p, gr in zip(params, gs):
            p.grad = None if gr is None else gr.clone()
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step_batch()
    return params, lrs, opt


def test_scaled_adam_matches_reference_run(emulated_adam):
    g = torch.load(os.path.join(GOLDEN, "ref_scaled_adam.pt"), weights_only=False)
    params, lrs, opt = _run(g)
    assert max(abs(a - b) for a, b in zip(lrs, g["lrs"])) < 1e-12
    for p, ref, (n, _) in zip(params, g["final"], g["shapes"]):
        assert rel_rms(p.detach(), ref) < 2e-5, n
    st = opt.state_dict()["state"]
    assert any("model_norms" in v for v in st.values())
    assert all(set(v) >= {"step", "exp_avg_sq", "delta"} for v in st.values())


def test_scaled_adam_checkpoint_resume_is_transparent(emulated_adam):
    g = torch.load(os.path.join(GOLDEN, "ref_scaled_adam.pt"), weights_only=False)
    a, _, _ = _run(g, upto=30)
    b, _, _ = _run(g, upto=30, resume_at=17)
    for x, y in zip(a, b):
        assert rel_rms(x.detach(), y.detach()) < 1e-6


def test_scaled_adam_many_tensors_vs_oracle(emulated_adam):
    from flow2gan_b200.optim import ScaledAdam
    gen = torch.Generator().manual_seed(11)
    shapes = [(64, 33, 3)] * 2 + [(257,)] * 4 + [()] * 5 + [(300, 1)] * 2 + [(5000, 3)]
    names = [f"p{i}" for i in range(len(shapes))]
    init = [torch.randn(s, generator=gen) * 0.2 for s in shapes]
    cpu = [p.clone() for p in init]
    mine = [torch.nn.Parameter(p.clone()) for p in init]
    ora = O.ScaledAdamOracle(names, cpu, lr=0.01, clipping_scale=2.0)
    opt = ScaledAdam(list(zip(names, mine)), lr=0.01, clipping_scale=2.0)
    for step in range(14):
        gs = [torch.randn(s, generator=gen) * (8.0 if step == 11 else 1.0) for s in shapes]
        ora.step(cpu, gs)
        for p, gr in zip(mine, gs):
            p.grad = gr
        opt.step()
    for a, b, n in zip(mine, cpu, names):
        assert rel_rms(a.detach(), b) < 2e-5, n


def test_state_dict_loads_into_reference_scaled_adam(emulated_adam):
    """Checkpoint interop (optim.py:84-101,304-339): a state_dict written here must put the clipping
    history where the REFERENCE reads it -- `tuples[0]`, the first batch in (str(dtype), *shape) key
    order -- and the reference optimizer must continue from it exactly like this one does."""
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "flow2gan")):
        pytest.skip("reference not mounted")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import import_reference
    import_reference()
    from flow2gan.optim import ScaledAdam as RefAdam
    from flow2gan_b200.optim import ScaledAdam
    gen = torch.Generator().manual_seed(3)
    # names chosen so that NAME order and KEY order of the batches differ: 'a...' has the largest shape key
    shapes = {"a_big": (40, 9), "b_vec": (12,), "c_scalar": (), "d_vec": (12,), "e_small": (3, 5)}
    init = {k: torch.randn(s, generator=gen) * 0.3 for k, s in shapes.items()}
    mine = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    opt = ScaledAdam(list(mine.items()), lr=0.02, clipping_scale=2.0, clipping_update_period=10)
    grads = [{k: torch.randn(s, generator=gen) for k, s in shapes.items()} for _ in range(16)]
    for gs in grads[:12]:
        for k, p in mine.items():
            p.grad = gs[k].clone()
        opt.step()
    import copy
    sd = copy.deepcopy(opt.state_dict())     # what torch.save / torch.load does (state_dict() itself returns live tensors)
    theirs = {k: torch.nn.Parameter(v.detach().clone()) for k, v in mine.items()}
    ref = RefAdam(list(theirs.items()), lr=0.02, clipping_scale=2.0, clipping_update_period=10)
    ref.load_state_dict(sd)
    # the reference's "first" batch: smallest (str(dtype), *shape) key
    keys = sorted({(str(p.dtype), *p.shape) for p in theirs.values()})
    first = next(p for p in theirs.values() if (str(p.dtype), *p.shape) == keys[0])
    st = ref.state[first]
    assert "model_norms" in st and "model_norm_threshold" in st and "num_clipped" in st
    assert float(st["model_norms"].abs().sum()) > 0
    for gs in grads[12:]:
        for k in shapes:
            mine[k].grad = gs[k].clone()
            theirs[k].grad = gs[k].clone()
        opt.step()
        ref.step()
    for k in shapes:
        assert rel_rms(mine[k].detach(), theirs[k].detach()) < 2e-5, k


def test_size_update_period_is_bounded():
    from flow2gan_b200.optim import ScaledAdam
    with pytest.raises(ValueError, match="size_update_period"):
        ScaledAdam([("p", torch.nn.Parameter(torch.zeros(4)))], lr=0.1, size_update_period=5)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less box")
def test_no_cpu_fallback():
    from flow2gan_b200.optim import ScaledAdam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
        ScaledAdam([("p", p)], lr=0.1).step()
