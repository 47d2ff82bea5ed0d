"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol that
include/flow2gan_b200.h declares, module trees / configs match the reference layout, and the
product path refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

from _cases import GOLDEN, ROOT
from oracle import flow2gan_oracle as O


def test_abi_library_loads_and_exports_all_declared_symbols():
    from flow2gan_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "flow2gan_b200.h")).read()
    declared = set(re.findall(r"\b(f2g_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/flow2gan_b200.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.f2g_abi_version() == 8


def test_state_dict_layout_matches_reference_spec():
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    for tag, name in (("24k", "mel_24k_base"), ("44k", "mel_44k_128band_512x_base")):
        g = torch.load(os.path.join(GOLDEN, f"ref_infer_{tag}.pt"), weights_only=False)
        m = MelAudioGenerator(**get_generator_config(name))
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == dict(g["sd_spec"])
        cfg, ocfg = get_generator_config(name), O.generator_config(name)
        for k, v in ocfg.items():
            assert cfg[k] == v, k
        # buffers equal the oracle's own construction
        assert torch.allclose(m.estimators[0].fft.window, O.hann(cfg["n_ffts"][0]), atol=2e-7)
        assert torch.allclose(m.loss_spec.fb, O.linear_fbanks(cfg["loss_n_fft"] // 2 + 1, 256,
                                                              cfg["sampling_rate"]))


def test_get_model_surface(tmp_path):
    import flow2gan_b200 as F
    with pytest.raises(ValueError):
        F.get_generator_config("nope")
    with pytest.raises(AssertionError):
        F.get_model("mel_24k_base", hf_model_name=None, checkpoint=None)
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**F.get_generator_config("mel_24k_base"))
    ck = tmp_path / "g.pt"
    torch.save({"model": {"module." + k: v for k, v in m.state_dict().items()}, "extra": 1}, ck)
    m2, cfg = F.get_model("mel_24k_base", checkpoint=str(ck))
    assert cfg.mel_hop_length == 256
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less box")
def test_no_cpu_fallback():
    import flow2gan_b200 as F
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**F.get_generator_config("mel_24k_base")).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m.infer(torch.zeros(1, 100, 8))
    from flow2gan_b200.modules import LogMelSpectrogram
    with pytest.raises((RuntimeError, AssertionError)):
        LogMelSpectrogram()(torch.zeros(1, 4000))


@pytest.mark.skipif(not os.path.isdir("/root/reference/flow2gan"), reason="reference not mounted")
def test_checkpoint_files_interoperate_with_reference(tmp_path):
    """Files written by flow2gan_b200.checkpoint.save_checkpoint load with the reference's
    load_checkpoint and vice versa (same top-level keys, DDP prefix handling, params passthrough,
    the reference's in-place fp32 downcast of model_avg)."""
    import copy
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    import_reference()
    import flow2gan.checkpoint as RC
    from flow2gan_b200 import checkpoint as MC
    from make_golden_datapath import toy_model
    m = toy_model()
    avg = copy.deepcopy(m).to(torch.float64)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    m.head.weight.sum().backward()
    opt.step()
    sched = torch.optim.lr_scheduler.StepLR(opt, 3)
    params = {"batch_idx_train": 77, "best_train_loss": 0.5}
    mine, ref = tmp_path / "mine.pt", tmp_path / "ref.pt"
    MC.save_checkpoint(mine, m, model_avg=avg, params=params, optimizer=opt, scheduler=sched, optimizer_disc=opt)
    assert next(avg.parameters()).dtype == torch.float32            # the reference's in-place downcast
    avg = avg.to(torch.float64)
    RC.save_checkpoint(ref, m, model_avg=avg, params=params, optimizer=opt, scheduler=sched, optimizer_disc=opt)
    a = torch.load(mine, weights_only=False)
    b = torch.load(ref, weights_only=False)
    assert list(a.keys()) == list(b.keys())
    for k in ("model", "model_avg"):
        assert list(a[k]) == list(b[k]) and all(torch.equal(a[k][n], b[k][n]) and a[k][n].dtype == b[k][n].dtype
                                                for n in a[k])
    assert a["batch_idx_train"] == 77 and a["optimizer"]["param_groups"] == b["optimizer"]["param_groups"]
    # cross loading, both directions, incl. a DDP-prefixed file
    for saver_file, loader in ((mine, RC.load_checkpoint), (ref, MC.load_checkpoint)):
        m2, avg2 = toy_model(), copy.deepcopy(toy_model()).to(torch.float64)
        with torch.no_grad():
            for p in m2.parameters():
                p.zero_()
        opt2 = torch.optim.SGD(m2.parameters(), lr=0.3, momentum=0.9)
        rest = loader(saver_file, m2, model_avg=avg2, optimizer=opt2)
        assert rest["batch_idx_train"] == 77 and "optimizer" not in rest and "model_avg" not in rest
        assert all(torch.equal(x, y) for x, y in zip(m.state_dict().values(), m2.state_dict().values()))
        assert opt2.param_groups[0]["lr"] == 0.1
    ddp = tmp_path / "ddp.pt"
    torch.save({"model": {"module." + k: v for k, v in m.state_dict().items()}}, ddp)
    m3 = toy_model()
    with torch.no_grad():
        m3.gain.fill_(5.0)
    MC.load_checkpoint(ddp, m3)
    assert all(torch.equal(x, y) for x, y in zip(m.state_dict().values(), m3.state_dict().values()))
    avg_keep = copy.deepcopy(m).to(torch.float64)
    MC.save_checkpoint(tmp_path / "keep.pt", m, model_avg=avg_keep, downcast_avg_in_place=False)
    assert next(avg_keep.parameters()).dtype == torch.float64
    assert torch.load(tmp_path / "keep.pt", weights_only=False)["model_avg"]["gain"].dtype == torch.float32


@pytest.mark.skipif(not os.path.isdir("/root/reference/flow2gan"), reason="reference not mounted")
def test_optimizer_host_logic_matches_reference_in_place():
    """ScaledAdam's four accepted parameter forms (optim.py:340-445) yield the reference's param_groups /
    parameters_names; Eden2 follows the reference's lr trajectory through warm-up and decay; the
    scheduler state_dict round-trips.  Host logic only: no optimizer step is taken (that needs the GPU)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    import_reference()
    import flow2gan.optim as RO
    import flow2gan_b200.optim as MO
    from make_golden_datapath import toy_model

    def forms(m):
        named = list(m.named_parameters())
        half = len(named) // 2
        return [
            [p for _, p in named],
            named,
            [{"params": [p for _, p in named[:half]], "lr": 0.01}, {"params": [p for _, p in named[half:]]}],
            [{"named_params": named[:half], "lr": 0.01}, {"named_params": named[half:], "lr": 0.002}],
        ]

    for fa, fb in zip(forms(toy_model()), forms(toy_model())):
        a = RO.ScaledAdam(fa, lr=0.045, clipping_scale=2.0)
        b = MO.ScaledAdam(fb, lr=0.045, clipping_scale=2.0)
        assert a.parameters_names == b.parameters_names
        assert a.show_dominant_parameters == b.show_dominant_parameters
        assert len(a.param_groups) == len(b.param_groups)
        for ga, gb in zip(a.param_groups, b.param_groups):
            assert [tuple(p.shape) for p in ga["params"]] == [tuple(p.shape) for p in gb["params"]]
            for k in ga:
                if k != "params":
                    assert ga[k] == gb[k], k
        sa = RO.Eden2(a, lr_batches=50, warmup_batches=20, warmup_start=0.1)
        sb = MO.Eden2(b, lr_batches=50, warmup_batches=20, warmup_start=0.1)
        for step in list(range(1, 30)) + [100, 1000]:
            sa.step_batch(step)
            sb.step_batch(step)
            assert sa.get_last_lr() == sb.get_last_lr(), step
            assert [g["lr"] for g in a.param_groups] == [g["lr"] for g in b.param_groups]
        sb2 = MO.Eden2(MO.ScaledAdam(forms(toy_model())[1], lr=0.045), lr_batches=50, warmup_batches=20)
        sb2.load_state_dict(sa.state_dict())
        assert sb2.batch == 1000 and sb2.state_dict() == sb.state_dict()
    with pytest.raises(ValueError):
        MO.ScaledAdam([], lr=0.1)


def test_gan_state_dict_layout_matches_reference_spec():
    """GAN(generator, ...) rebuilds the reference's module tree: same 666 state_dict keys and shapes as
    the reference GAN the golden file was generated from (checkpoints load with strict=False, so a
    silent key mismatch would just skip weights)."""
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    g = torch.load(os.path.join(GOLDEN, "ref_gan_24k.pt"), weights_only=False)
    gan = GAN(MelAudioGenerator(**get_generator_config("mel_24k_base")), **get_gan_config("gan_multi_scale_mel_recon"))
    mine = [(k, tuple(v.shape)) for k, v in gan.state_dict().items()]
    assert mine == [(k, tuple(s)) for k, s in g["sd_spec"]]
    n_gen = sum(p.numel() for p in gan.generator.parameters())
    n_disc = sum(p.numel() for p in gan.discriminator.parameters())
    assert (n_gen, n_disc) == (78949542, 42503752)


@pytest.mark.skipif(not os.path.isdir("/root/reference/flow2gan"), reason="reference not mounted")
def test_configs_and_filterbanks_match_reference_in_place():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    import_reference()
    import flow2gan.models.config as RC
    import flow2gan_b200.config as MC
    for name in ("mel_24k_base", "mel_44k_128band_512x_base"):
        assert dict(RC.get_generator_config(name)) == dict(MC.get_generator_config(name)), name
    assert dict(RC.get_gan_config("gan_multi_scale_mel_recon")) == dict(MC.get_gan_config("gan_multi_scale_mel_recon"))
    assert RC.HF_MODEL_NAMES == MC.HF_MODEL_NAMES and RC.HF_REPO == MC.HF_REPO
    for bad in ("nope", ""):
        with pytest.raises(ValueError):
            MC.get_generator_config(bad)
    import torchaudio
    from flow2gan_b200.modules import linear_fbanks, melscale_fbanks
    for n_fft, n_mels in zip((32, 64, 128, 256, 512, 1024, 2048), (5, 10, 20, 40, 80, 160, 320)):
        ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, 12000.0, n_mels, 24000, norm=None, mel_scale="htk")
        assert torch.equal(melscale_fbanks(n_fft // 2 + 1, n_mels, 24000), ref), n_fft
    ref = torchaudio.functional.linear_fbanks(513, 0.0, 12000.0, 256, 24000)
    assert torch.equal(linear_fbanks(513, 256, 24000), ref)


def test_branch_dropout_draws_follow_reference_order():
    """generator.py:145-162: randint (branch to drop) then rand (apply with prob. p), mask rescaled by
    nb / (nb - 1); ours additionally folds the branch mean (1 / nb) into the weights."""
    from types import SimpleNamespace
    from flow2gan_b200.train import _branch_dropout_weight
    m = SimpleNamespace(num_branches=3, training=True, branch_dropout=0.5)
    torch.manual_seed(11)
    w = _branch_dropout_weight(m, 64, torch.device("cpu"))
    torch.manual_seed(11)
    idx = torch.randint(0, 3, (64,))
    mask = torch.ones(64, 3)
    mask[torch.arange(64), idx] = 0.0
    mask = mask * (3 / 2)
    want = torch.where(torch.rand(64, 1) < 0.5, mask, torch.ones_like(mask))
    assert torch.equal(w, want / 3)
    assert 10 < int((w == 0).sum()) < 54 and torch.allclose(w.sum(1), torch.ones(64))
    m.training = False
    assert _branch_dropout_weight(m, 64, torch.device("cpu")) is None


@pytest.mark.skipif(not os.path.isdir("/root/reference/flow2gan"), reason="reference not mounted")
def test_limit_param_value_flip_matches_reference_in_place():
    """train._limit_flip / _limit_apply == LimitParamValue.backward (modules.py:236-256) on random
    gradients with parameters below, inside and above the range; the graph-mode device flag gives the
    same result as the eager Python flag."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    import_reference()
    from flow2gan.models.modules import LimitParamValue
    from flow2gan_b200.train import _limit_apply, _limit_flip
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(4000, generator=g) * 1.5)
    grad = torch.randn(4000, generator=g)
    grad[::7] = 0.0
    for lo, hi in ((0.5, 1.0), (-1.5, 1.5)):
        xr = x.clone().requires_grad_(True)
        LimitParamValue.apply(xr, lo, hi).backward(grad.clone())
        got = _limit_flip(grad.clone(), x, lo, hi)
        assert torch.equal(got, xr.grad)
        assert torch.equal(_limit_apply(True, grad, x, lo, hi), got)
        assert torch.equal(_limit_apply(False, grad, x, lo, hi), grad)
        assert torch.equal(_limit_apply(torch.tensor(1.0), grad, x, lo, hi), got)
        assert torch.equal(_limit_apply(torch.tensor(0.0), grad, x, lo, hi), grad)
        assert int((got != grad).sum()) > 100


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line with the contract's keys, timed on the
    UNMODIFIED reference staged under baseline/_ref (kind "reference"); needs no GPU."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "baseline", "_ref", "flow2gan")):
        pytest.skip("reference not staged (tools/stage_reference.sh needs /root/reference)")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["warmup"] >= 3
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("mel_24k_base 1-step inference") and d["value"] > 0
