"""Dry run of the WHOLE inference path on a GPU-less box: the product's Python host layer (weight
packing, launch sequence, Euler sampler: flow2gan_b200/engine.py, generator.py) drives the SAME SIMT
kernel sources under the host emulation (blocks.cu, spectral.cu) plus a plain host restatement of the
tensor-core contraction contract (tests/emul_gemm.cpp), and the result is compared with outputs of
the reference itself (tests/golden/ref_infer_24k.pt) at the parity gate of BASELINE.json (1e-3
rel-RMS).  This checks everything except the tcgen05 / TMA kernels themselves, which only the GPU
suite can run."""
import os

import pytest
import torch

import _emul
from _cases import GOLDEN, rel_rms
from _synth import synth_state_dict

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


@pytest.fixture
def L(monkeypatch):
    lib = _emul.native_fixture(monkeypatch)
    import flow2gan_b200.engine as E
    from flow2gan_b200.generator import BaseAudioGenerator
    monkeypatch.setattr(E, "FORK_COND", False)                       # no side stream on the CPU
    monkeypatch.setattr(BaseAudioGenerator, "_require_cuda", lambda self: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)    # eager launches, no CUDA graph
    return lib


def tf32_round(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("epi", ["plain", "bias_prelu_round", "res", "rowscale_leaky", "gate_acc", "silu", "f16_mlp"])
def test_host_gemm_restatement_follows_the_contract(L, epi):
    """tests/emul_gemm.cpp against torch for every epilogue of include/flow2gan_b200.h::F2GGemm (the
    same cases tests/test_kernels_gpu.py runs on the tcgen05 kernel)."""
    g = torch.Generator().manual_seed(3)
    M, N, K = 70, 52, 96
    A, Bm = tf32_round(torch.randn(M, K, generator=g)), tf32_round(torch.randn(N, K, generator=g))
    ref = A.double() @ Bm.double().t()
    bias, slope = torch.randn(N, generator=g), torch.rand(N, generator=g) * 0.5
    res, rsc = torch.randn(M, N, generator=g), torch.rand(N, generator=g) + 0.5
    rows = (torch.rand(M, generator=g) > 0.3).float()
    gate = torch.randn(M, N, generator=g)
    ldc = N + 4
    C0 = torch.randn(M, ldc, generator=g)
    Cd = C0.clone()
    kw, a, b, tol = {}, A, Bm, 1e-6
    if epi == "plain":
        exp = ref
    elif epi == "bias_prelu_round":
        z = ref + bias.double()
        exp = tf32_round(torch.where(z > 0, z, z * slope.double()).float()).double()
        kw = dict(bias=bias.data_ptr(), slope=slope.data_ptr(), act=L.ACT_PRELU, round_tf32=1)
        tol = 3e-4
    elif epi == "res":
        exp = ref + bias.double() + rsc.double() * res.double()
        kw = dict(bias=bias.data_ptr(), res=res.data_ptr(), ld_res=N, res_scale=rsc.data_ptr())
    elif epi == "rowscale_leaky":
        z = ref * 0.5 + bias.double()
        exp = torch.where(z > 0, z, z * 0.1) * rows.double()[:, None]
        kw = dict(bias=bias.data_ptr(), act=L.ACT_LEAKY, leaky=0.1, alpha=0.5, row_scale=rows.data_ptr())
    elif epi == "gate_acc":
        exp = ref * torch.where(gate > 0, 1.0, slope.double()[None].expand(M, N)) + C0[:, :N].double()
        kw = dict(slope=slope.data_ptr(), gate=gate.data_ptr(), ld_gate=N, accumulate=1)
    elif epi == "silu":
        z = ref + bias.double()
        exp = z * torch.sigmoid(z)
        kw = dict(bias=bias.data_ptr(), act=L.ACT_SILU)
    elif epi == "f16_mlp":
        a, b = A.half(), Bm.half()
        z = a.double() @ b.double().t() + bias.double()
        exp = torch.where(z > 0, z, z * slope.double()).half().double()
        Cd = torch.zeros(M, 56, dtype=torch.float16)
        ldc = 56
        kw = dict(bias=bias.data_ptr(), slope=slope.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1)
        tol = 1e-3
    L.gemm_group([L.gemm_desc(a.data_ptr(), b.data_ptr(), Cd.data_ptr(), M, N, K, K, K, ldc, **kw)])
    assert rel_rms(Cd[:, :N].double(), exp) < tol
    if epi != "f16_mlp":
        assert torch.equal(Cd[:, N:], C0[:, N:])


def _model(g):
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**get_generator_config(g["model_name"]))
    m.load_state_dict(synth_state_dict(g["sd_spec"], g["sd_seed"]), strict=False)
    return m.eval()


def test_model_infer_dry_run_matches_reference_golden(L):
    """One masked 1-step `model.infer` (audio_lens given: frame masks, length = lens.max()) by default --
    about a minute of emulation; F2G_SLOW_TESTS=1 adds the unmasked 1- and 2-step and the clamped runs
    (8 minutes in total)."""
    g = torch.load(os.path.join(GOLDEN, "ref_infer_24k.pt"), weights_only=False)
    m = _model(g)
    errs = {}
    with torch.no_grad():
        lens = g["lens"]
        out = m.infer(g["mel"], audio_lens=lens, n_timesteps=1, noise=g["noise"][:, : int(lens.max())])
        assert out.shape == g["audio_lens_n1"].shape
        errs["lens"] = rel_rms(out, g["audio_lens_n1"])
        if os.environ.get("F2G_SLOW_TESTS") == "1":
            for n in (1, 2):
                out = m.infer(g["mel"], n_timesteps=n, noise=g["noise"])
                errs[n] = rel_rms(out, g[f"audio_n{n}"])
            out = m.infer(g["mel"], n_timesteps=2, clamp_pred=True, noise=g["noise"] * 30)
            errs["clamp"] = rel_rms(out, g["audio_n2_clamp"])
    print("dry-run rel-RMS vs reference:", errs)
    assert all(v < 1e-3 for v in errs.values()), errs


def test_ragged_tiny_batch_vs_oracle(L):
    """Edge shapes the golden files do not hold: 4 mel frames, ragged lengths that are no multiple of
    the hop (1023 and 529 samples), 2 Euler steps -- against the CPU oracle."""
    from _cases import mel_input, noise_input
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from oracle import flow2gan_oracle as O
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 7)
    m.load_state_dict(sd, strict=False)
    m.eval()
    mel, noise, lens = mel_input(2, 100, 4, seed=4), noise_input(2, 1023, seed=5), torch.tensor([1023, 529])
    with torch.no_grad():
        out = m.infer(mel, audio_lens=lens, n_timesteps=2, noise=noise)
        ref = O.generator_infer(sd, O.generator_config("mel_24k_base"), mel, noise, lens, 2, False)
    assert out.shape == (2, 1023) and rel_rms(out, ref) < 1e-3


@pytest.mark.skipif(os.environ.get("F2G_SLOW_TESTS") != "1", reason="set F2G_SLOW_TESTS=1 (minutes of emulation)")
def test_streaming_dry_run_matches_oracle_chunk_by_chunk(L):
    """flow2gan_b200.streaming (bin/infer_dir.py:126-168) on the emulated kernels: chunked synthesis with
    24 frames of context equals the oracle chunk by chunk; stacking the interior chunks along the batch
    axis gives the same numbers."""
    from _cases import mel_input, noise_input
    from flow2gan_b200 import AttributeDict, get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.streaming import chunk_plan, streaming_infer_audio
    from oracle import flow2gan_oracle as O
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99)
    m.load_state_dict(sd, strict=False)
    m.eval()
    B, frames, chunk, hop = 1, 70, 16, 256
    mel = mel_input(B, 100, frames, seed=5)
    params = AttributeDict(n_timesteps=1, chunk_size=chunk)
    plan = chunk_plan(frames, chunk, hop)
    noises = [noise_input(B, (f1 - f0) * hop, seed=40 + i) for i, (f0, f1, _, _) in enumerate(plan)]
    got = streaming_infer_audio(params, m, None, cond=mel, noise_fn=lambda i, shape: noises[i])
    assert got.shape == (B, frames * hop)
    cfg = O.generator_config("mel_24k_base")
    ref = []
    with torch.no_grad():
        for (f0, f1, lp, rp), nz in zip(plan, noises):
            a = O.generator_infer(sd, cfg, mel[:, :, f0:f1], nz, None, 1, True)
            ref.append(a[:, lp: a.size(1) - rp])
    assert rel_rms(got, torch.cat(ref, -1)) < 1e-3
    got_b = streaming_infer_audio(params, m, None, cond=mel, batch_chunks=True, noise_fn=lambda i, shape: noises[i])
    assert rel_rms(got_b, got) < 1e-5


def test_cached_time_path_is_bit_identical(L, monkeypatch):
    """F2G_CACHE_TIME (default on): per-step time-scale vectors computed once per (plan, N, weight
    version) -- the 2-step output must equal the uncached run bit for bit (same kernels, same inputs,
    fewer launches); after an in-place weight refresh the cache is recomputed; under an outer stream
    capture (the trainer's phase graphs) it is recomputed on every call."""
    import flow2gan_b200.engine as E
    g = torch.load(os.path.join(GOLDEN, "ref_infer_24k.pt"), weights_only=False)
    outs, counts = [], []
    for flag in (False, True):
        monkeypatch.setattr(E, "CACHE_TIME", flag)
        monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
        m = _model(g)
        B, _, Fm = g["mel"].shape
        with torch.no_grad():
            plan = m.plan(B, Fm, g["noise"].shape[-1], False)
            plan.infer(g["mel"], g["noise"], None, 2, False, use_graph=False)      # builds the cache
            L.COUNT = 0
            outs.append(plan.infer(g["mel"], g["noise"], None, 2, False, use_graph=False).clone())
            counts.append(L.COUNT)
            if flag:
                ver = m._packed.version
                m.estimators[0].decoder.time_mlp[0].bias.data.add_(0.25)           # weights move ...
                torch.autograd.graph.increment_version(list(m.parameters()))
                assert m.plan(B, Fm, g["noise"].shape[-1], False) is plan           # ... the plan survives
                assert m._packed.version == ver + 1
                L.COUNT = 0
                moved = plan.infer(g["mel"], g["noise"], None, 2, False, use_graph=False)
                assert L.COUNT == counts[1] + 2 * 4                                 # cache rebuilt once
                assert not torch.equal(moved, outs[1])
                monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)
                L.COUNT = 0
                again = plan.infer(g["mel"], g["noise"], None, 2, False, use_graph=False)
                assert L.COUNT == counts[1] + 2 * 4 and torch.equal(again, moved)   # outer capture: always
    assert torch.equal(outs[0], outs[1])
    assert counts[1] == counts[0] - 2 * 4, counts                     # 4 launches per ODE step left the sequence
    assert rel_rms(outs[1], g["audio_n2"]) < 1e-3
