"""Unit parity tests of the CUDA kernels (through the C ABI / ctypes) against the CPU oracle.
Tolerances: fp32 SIMT kernels 1e-5 rel-RMS; TF32 tensor-core GEMM with TF32-exact inputs 2e-6."""
import math

import pytest
import torch

from _cases import audio_input, rel_rms
from oracle import flow2gan_oracle as O

pytestmark = pytest.mark.gpu


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 on CPU: round-to-nearest (ties away) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


@pytest.fixture(scope="module")
def L():
    from flow2gan_b200 import _lib
    _lib.lib()
    return _lib


def _gemm_case(L, M, N, K, bn, a_mn=0, b_mn=0, epi="plain", lda_pad=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = tf32_round(torch.randn(M, K, generator=g))
    Bm = tf32_round(torch.randn(N, K, generator=g))
    lda = (K if not a_mn else M) + lda_pad
    ldb = (K if not b_mn else N) + lda_pad
    lda, ldb = (lda + 3) // 4 * 4, (ldb + 3) // 4 * 4
    if a_mn:
        Ad = torch.zeros(K, lda); Ad[:, :M] = A.t()
    else:
        Ad = torch.zeros(M, lda); Ad[:, :K] = A
    if b_mn:
        Bd = torch.zeros(K, ldb); Bd[:, :N] = Bm.t()
    else:
        Bd = torch.zeros(N, ldb); Bd[:, :K] = Bm
    ldc = (N + 3) // 4 * 4 + 4
    ref = (A.double() @ Bm.double().t())
    bias = torch.randn(N, generator=g)
    slope = torch.rand(N, generator=g) * 0.5
    res = torch.randn(M, N, generator=g)
    rsc = torch.rand(N, generator=g) + 0.5
    rows = (torch.rand(M, generator=g) > 0.3).float()
    gate = torch.randn(M, N, generator=g)
    C0 = torch.randn(M, ldc, generator=g)
    kw = {}
    dev = "cuda"
    keep = []

    def d(t):
        t = t.to(dev).contiguous(); keep.append(t); return t.data_ptr()

    if epi == "plain":
        exp = ref
    elif epi == "bias_prelu_round":
        z = ref + bias.double()
        exp = tf32_round(torch.where(z > 0, z, z * slope.double()).float()).double()
        kw = dict(bias=d(bias), slope=d(slope), act=L.ACT_PRELU, round_tf32=1)
    elif epi == "res":
        exp = ref + bias.double() + rsc.double() * res.double()
        kw = dict(bias=d(bias), res=d(res), ld_res=N, res_scale=d(rsc))
    elif epi == "rowscale_leaky":
        z = ref * 0.5 + bias.double()
        exp = torch.where(z > 0, z, z * 0.1) * rows.double()[:, None]
        kw = dict(bias=d(bias), act=L.ACT_LEAKY, leaky=0.1, alpha=0.5, row_scale=d(rows))
    elif epi == "gate_acc":
        exp = ref * torch.where(gate > 0, 1.0, slope.double()[None].expand(M, N)) + C0[:, :N].double()
        kw = dict(slope=d(slope), gate=d(gate), ld_gate=N, accumulate=1)
    elif epi == "silu":
        z = ref + bias.double()
        exp = z * torch.sigmoid(z)
        kw = dict(bias=d(bias), act=L.ACT_SILU)
    Cd = C0.clone().to(dev)
    Ag, Bg = Ad.to(dev), Bd.to(dev)
    L.gemm_group([L.gemm_desc(Ag.data_ptr(), Bg.data_ptr(), Cd.data_ptr(), M, N, K, lda, ldb, ldc,
                              bn=bn, a_mn=a_mn, b_mn=b_mn, **kw)])
    torch.cuda.synchronize()
    got = Cd.cpu()
    assert torch.equal(got[:, N:], C0[:, N:]), "GEMM wrote outside its N columns"
    tol = 3e-4 if epi in ("bias_prelu_round",) else 1e-5   # fp32 accumulation over K<=2304
    err = rel_rms(got[:, :N], exp)
    assert err < tol, (M, N, K, bn, a_mn, b_mn, epi, err)


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 32, 128), (128, 64, 64, 64), (300, 256, 96, 128), (300, 256, 96, 256),
    (1520, 2304, 768, 128), (1520, 768, 2304, 64), (333, 130, 130, 128), (97, 514, 768, 128),
    (6032, 1152, 384, 256),
])
def test_gemm_tn_shapes(L, M, N, K, bn):
    _gemm_case(L, M, N, K, bn)


@pytest.mark.parametrize("epi", ["bias_prelu_round", "res", "rowscale_leaky", "gate_acc", "silu"])
def test_gemm_epilogues(L, epi):
    _gemm_case(L, 421, 384, 160, 128, epi=epi)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,bn", [(256, 128, 64, 128), (300, 200, 100, 128), (768, 2304, 1520, 256),
                                      (200, 96, 333, 64)])
def test_gemm_mn_major(L, a_mn, b_mn, M, N, K, bn):
    _gemm_case(L, M, N, K, bn, a_mn=a_mn, b_mn=b_mn)


def test_gemm_grouped(L):
    g = torch.Generator().manual_seed(3)
    descs, checks = [], []
    for (M, N, K) in ((1520, 768, 514), (3024, 512, 258), (6032, 384, 130), (17, 384, 40)):
        ld = (K + 3) // 4 * 4
        A = torch.zeros(M, ld); A[:, :K] = tf32_round(torch.randn(M, K, generator=g))
        Bm = torch.zeros(N, ld); Bm[:, :K] = tf32_round(torch.randn(N, K, generator=g))
        Ag, Bg, Cg = A.cuda(), Bm.cuda(), torch.zeros(M, N).cuda()
        descs.append(L.gemm_desc(Ag.data_ptr(), Bg.data_ptr(), Cg.data_ptr(), M, N, K, ld, ld, N))
        checks.append((Ag, Bg, Cg, A[:, :K].double() @ Bm[:, :K].double().t()))
    L.gemm_group(descs)
    torch.cuda.synchronize()
    for Ag, Bg, Cg, ref in checks:
        assert rel_rms(Cg.cpu(), ref) < 3e-6


@pytest.mark.parametrize("M,N,K,epi", [
    (1520, 2304, 768, "prelu_f16"), (6032, 1152, 384, "prelu_f16"), (300, 256, 96, "prelu_f16"),
    (1505, 1536, 512, "prelu_f16"), (333, 160, 72, "prelu_f16"),
    (333, 120, 72, "prelu_f16"), (300, 248, 96, "prelu_f16"),     # TMA-store path, last box clipped in N and M
    (300, 250, 96, "prelu_f16"),                                  # N % 8 != 0: plain-store path (16-byte clip granule)
    (1520, 768, 2304, "res"), (6032, 384, 1152, "res"), (97, 514, 776, "plain"), (421, 384, 160, "res_round"),
])
def test_gemm_f16_operands(L, M, N, K, epi):
    """kind::f16 GEMM (fp16 operands exact in the inputs, fp32 accumulate) vs float64 matmul:
    the ConvNeXt-block contractions of the inference engine (modules.py:443-451)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).half()
    Bm = (torch.randn(N, K, generator=g) * 0.05).half()
    lda = (K + 7) // 8 * 8 + 8
    Ad = torch.zeros(M, lda, dtype=torch.float16); Ad[:, :K] = A
    Bd = torch.zeros(N, lda, dtype=torch.float16); Bd[:, :K] = Bm
    ref = A.double() @ Bm.double().t()
    bias = torch.randn(N, generator=g)
    slope = torch.rand(N, generator=g) * 0.5
    res = torch.randn(M, N, generator=g)
    rsc = torch.rand(N, generator=g) + 0.5
    keep = []

    def d(t):
        t = t.cuda().contiguous(); keep.append(t); return t.data_ptr()

    Ag, Bg = Ad.cuda(), Bd.cuda()
    if epi == "prelu_f16":
        ldc = (N + 7) // 8 * 8 + 8
        C0 = torch.randn(M, ldc, generator=g).half()
        Cd = C0.clone().cuda()
        z = ref + bias.double()
        exp = torch.where(z > 0, z, z * slope.double())
        L.gemm_group([L.gemm_desc(Ag.data_ptr(), Bg.data_ptr(), Cd.data_ptr(), M, N, K, lda, lda, ldc,
                                  bias=d(bias), slope=d(slope), act=L.ACT_PRELU, ab_f16=1, c_f16=1)])
        tol = 3e-4                                       # fp16 output rounding: 2^-11 / sqrt(3)
    else:
        ldc = N + 4
        C0 = torch.randn(M, ldc, generator=g)
        Cd = C0.clone().cuda()
        if epi == "plain":
            exp, kw = ref, {}
        else:
            Cd[:, :N] = res.cuda()                       # in-place residual stream, as the engine does
            exp = ref + bias.double() + rsc.double() * res.double()
            kw = dict(bias=d(bias), res=Cd.data_ptr(), ld_res=ldc, res_scale=d(rsc),
                      round_tf32=int(epi == "res_round"))
        L.gemm_group([L.gemm_desc(Ag.data_ptr(), Bg.data_ptr(), Cd.data_ptr(), M, N, K, lda, lda, ldc,
                                  ab_f16=1, **kw)])
        tol = 3e-4 if epi == "res_round" else 1e-5
    torch.cuda.synchronize()
    got = Cd.cpu()
    assert torch.equal(got[:, N:], C0[:, N:]), "GEMM wrote outside its N columns"
    err = rel_rms(got[:, :N].float(), exp)
    assert err < tol, (M, N, K, epi, err)


def test_gemm_f16_group_matches_tf32_group(L):
    """Three problems in one launch (the three branches), fp16 vs TF32 operands holding the same
    values: the products are identical, only the accumulation order may differ."""
    g = torch.Generator().manual_seed(11)
    d16, d32, outs = [], [], []
    for (M, N, K) in ((1520, 768, 2304), (3024, 512, 1536), (6032, 384, 1152)):
        A = torch.randn(M, K, generator=g).half()
        Bm = (torch.randn(N, K, generator=g) * 0.03).half()
        Ah, Bh, A32, B32 = A.cuda(), Bm.cuda(), A.float().cuda(), Bm.float().cuda()
        C16, C32 = torch.zeros(M, N).cuda(), torch.zeros(M, N).cuda()
        d16.append(L.gemm_desc(Ah.data_ptr(), Bh.data_ptr(), C16.data_ptr(), M, N, K, K, K, N, ab_f16=1))
        d32.append(L.gemm_desc(A32.data_ptr(), B32.data_ptr(), C32.data_ptr(), M, N, K, K, K, N))
        outs.append((C16, C32, Ah, Bh, A32, B32))
    L.gemm_group(d16)
    L.gemm_group(d32)
    torch.cuda.synchronize()
    for C16, C32, *_ in outs:
        assert rel_rms(C16, C32) < 1e-5


@pytest.mark.parametrize("shapes", [
    ((1520, 768), (3024, 512), (6032, 384)),          # the three branches of one block (bench shape)
    ((1505, 512),),                                   # CondEncoder block (+ the zero row)
    ((300, 384), (77, 512)),                          # fewer tiles than CTA pairs
])
def test_gemm_chained_mlp_matches_two_launches(L, shapes):
    """pwconv1 -> PReLU -> pwconv2 (+ residual) as ONE chained launch (per-row-tile counters) must
    give bit-identical results to the two-launch sequence (modules.py:486-493), repeatedly."""
    g = torch.Generator().manual_seed(len(shapes) * 100 + shapes[0][0])
    keep, chained, g1s, g2s = [], [], [], []
    n_cnt = sum((M + 255) // 256 for M, _ in shapes)
    cnt = torch.zeros(n_cnt, dtype=torch.int32, device="cuda")
    off = 0
    for M, C in shapes:
        H = 3 * C
        a = torch.randn(M, C, generator=g).half().cuda()
        W1 = (torch.randn(H, C, generator=g) * 0.04).half().cuda()
        W2 = (torch.randn(C, H, generator=g) * 0.03).half().cuda()
        b1, sl = torch.randn(H, generator=g).cuda(), (torch.rand(H, generator=g) * 0.5).cuda()
        b2, rs = torch.randn(C, generator=g).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
        x0 = torch.randn(M, C, generator=g).cuda()
        xa, xb = x0.clone(), x0.clone()
        ha = torch.zeros(M, H, dtype=torch.float16, device="cuda")
        hb = torch.zeros(M, H, dtype=torch.float16, device="cuda")
        keep += [a, W1, W2, b1, sl, b2, rs, x0, xa, xb, ha, hb]
        cp = cnt.data_ptr() + 4 * off
        off += (M + 255) // 256

        def d1(h, done=None):
            return L.gemm_desc(a.data_ptr(), W1.data_ptr(), h.data_ptr(), M, H, C, C, C, H, bias=b1.data_ptr(),
                               slope=sl.data_ptr(), act=L.ACT_PRELU, ab_f16=1, c_f16=1, done_counter=done)

        def d2(h, x, wait=None):
            return L.gemm_desc(h.data_ptr(), W2.data_ptr(), x.data_ptr(), M, C, H, H, H, C, bias=b2.data_ptr(),
                               res=x.data_ptr(), ld_res=C, res_scale=rs.data_ptr(), ab_f16=1, wait_counter=wait)
        g1s.append(d1(ha)); g2s.append(d2(ha, xa))
        chained += [(d1(hb, cp), d2(hb, xb, cp), x0, xa, xb, ha, hb)]
    L.gemm_group(g1s)
    L.gemm_group(g2s)
    torch.cuda.synchronize()
    descs = [c[0] for c in chained] + [c[1] for c in chained]
    for it in range(12):
        cnt.zero_()
        for _, _, x0, xa, xb, ha, hb in chained:
            xb.copy_(x0)
            hb.fill_(float("nan"))                  # a consumer that runs ahead of its producer reads NaN
        L.gemm_group(descs)
        torch.cuda.synchronize()
        for _, _, x0, xa, xb, ha, hb in chained:
            assert torch.equal(hb, ha), it
            assert torch.equal(xb, xa), (it, float((xb - xa).abs().max()))
    # sanity of the reference sequence itself against float64
    d1_, d2_, x0, xa, xb, ha, hb = chained[0]
    M, C = shapes[0]
    a, W1, W2, b1, sl, b2, rs = keep[:7]
    z = a.double().cpu() @ W1.double().cpu().t() + b1.double().cpu()
    h = torch.where(z > 0, z, z * sl.double().cpu()).half().double()
    exp = h @ W2.double().cpu().t() + b2.double().cpu() + rs.double().cpu() * x0.double().cpu()
    assert rel_rms(xa.cpu(), exp) < 3e-4


@pytest.mark.parametrize("rows,cols,act", [(5000, 32, "leaky"), (777, 48, "prelu"), (1505, 1152, "prelu"),
                                           (300, 30, "leaky"), (1000, 64, "none"), (515, 1536, "silu")])
def test_act_bwd_fused_reductions(L, rows, cols, act):
    """dz = dh * act'(z) with the fused bias / slope gradient column sums (autograd of PReLU,
    LeakyReLU, SiLU after a conv: modules.py:486-489,570-572, discriminators.py:99-104), in place,
    padded leading dimensions; vector (cols % 4 == 0) and scalar kernels."""
    g = torch.Generator().manual_seed(rows + cols)
    ld = (cols + 3) // 4 * 4 + 4
    dh = torch.randn(rows, ld, generator=g)
    z = torch.randn(rows, ld, generator=g)
    slope = torch.rand(cols, generator=g) * 0.5
    code = {"leaky": L.ACT_LEAKY, "prelu": L.ACT_PRELU, "none": L.ACT_NONE, "silu": L.ACT_SILU}[act]
    d, zz = dh[:, :cols].double(), z[:, :cols].double()
    if act == "leaky":
        exp = torch.where(zz > 0, d, d * 0.1)
    elif act == "prelu":
        exp = torch.where(zz > 0, d, d * slope.double())
    elif act == "silu":
        sg = torch.sigmoid(zz)
        exp = d * sg * (1 + zz * (1 - sg))
    else:
        exp = d
    dzg = dh.clone().cuda()
    gb = torch.zeros(cols + 4, device="cuda")
    gs = torch.zeros(cols + 4, device="cuda")
    L.act_bwd(dzg, ld, z.cuda() if act != "none" else None, ld, slope.cuda() if act == "prelu" else None, 0.1, code,
              rows, cols, dzg, ld, gb, gs if act in ("prelu", "leaky") else None)
    torch.cuda.synchronize()
    got = dzg.cpu()
    assert torch.equal(got[:, cols:], dh[:, cols:]), "wrote outside its columns"
    assert rel_rms(got[:, :cols], exp) < 1e-6
    assert rel_rms(gb[:cols].cpu(), exp.sum(0)) < 1e-5 and float(gb[cols:].abs().max()) == 0.0
    if act in ("prelu", "leaky"):
        assert rel_rms(gs[:cols].cpu(), (d * zz.clamp(max=0)).sum(0)) < 1e-5


@pytest.mark.parametrize("n_fft,hop", [(128, 64), (256, 128), (512, 256), (1024, 512), (32, 8), (2048, 512)])
def test_stft_packed(L, n_fft, hop):
    B, T = 3, 6144
    x = audio_input(B, T, seed=n_fft)
    ref = O.stft_packed(x, n_fft, hop)                 # (B, n+2, F)
    F = ref.shape[-1]
    ld = n_fft + 8
    out = torch.full((B * F, ld), 7.0, device="cuda")
    L.stft(x.cuda(), B, T, T, n_fft, hop, L.SPEC_PACKED, out, ld)
    got = out.cpu().view(B, F, ld)[:, :, : n_fft + 2].transpose(1, 2)
    assert rel_rms(got, ref) < 2e-6
    assert float(out[:, n_fft + 2:].abs().max()) == 0.0


@pytest.mark.parametrize("n_ffts,T", [((512, 256, 128), 6144), ((1024, 512, 256), 7168), ((256,), 1000)])
def test_warp_fft_group_vs_oracle(L, n_ffts, T):
    """Grouped warp-shuffle STFT / inverse (one launch for the branch resolutions,
    modules.py:699-719) against the oracle's torch.stft / istft restatement; hop = n/2, reflect
    edges on both sides, TF32 rounding off."""
    B = 3
    x = audio_input(B, T, seed=T)
    xg = x.cuda()
    outs, probs = [], []
    for n in n_ffts:
        hop = n // 2
        F = 1 + T // hop
        ld = n + 4
        o = torch.full((B * F, ld), 7.0, device="cuda")
        outs.append((n, hop, F, ld, o))
        probs.append((xg, o, n, hop, F, B * F, T, ld))
    L.stft_group(probs, B, T, round_tf32=0)
    frs, iprobs = [], []
    for n, hop, F, ld, o in outs:
        ref = O.stft_packed(x, n, hop)
        got = o.cpu().view(B, F, ld)[:, :, : n + 2].transpose(1, 2)
        assert rel_rms(got, ref) < 2e-6, n
        assert float(o[:, n + 2:].abs().max()) == 0.0
        fr = torch.empty(B * F, n, device="cuda")
        frs.append(fr)
        iprobs.append((o, fr, n, 0, 0, B * F, ld, n))
    L.irfft_group(iprobs)
    for (n, hop, F, ld, o), fr in zip(outs, frs):
        ref_fr = torch.empty(B * F, n, device="cuda")
        L.irfft_frames(o, B * F, ld, n, ref_fr)               # shared-memory Stockham kernel
        assert rel_rms(fr, ref_fr) < 2e-6, n
        out = torch.empty(B, T, device="cuda")
        L.ola_combine([fr], [n], [hop], [F], None, None, out, B, T, False, 0.0, 0.0, False)
        keep = hop * (F - 1)
        assert rel_rms(out.cpu()[:, :keep], x[:, :keep]) < 3e-6      # STFT -> iSTFT round trip


def test_logmel_fused_vs_oracle_and_fixture():
    import os
    from _cases import GOLDEN
    from flow2gan_b200.modules import LogMelSpectrogram
    g = torch.load(os.path.join(GOLDEN, "mel_24k_short.pt"), weights_only=False)
    wav = (g["pcm_int16"].float()[None] / 32768.0).cuda()
    m = LogMelSpectrogram(24000, 1024, 256, 100).cuda()
    got = m(wav).cpu()
    assert got.shape == g["mel"].shape
    assert rel_rms(got, g["mel"]) < 1e-5                       # the reference's own fixture
    x = audio_input(2, 12000, seed=9)
    assert rel_rms(m(x.cuda()).cpu(), O.log_mel(x)) < 1e-5
    m44 = LogMelSpectrogram(44100, 2048, 512, 128).cuda()
    assert rel_rms(m44(x.cuda()).cpu(), O.log_mel(x, 44100, 2048, 512, 128)) < 1e-5


@pytest.mark.parametrize("n_fft", [128, 256, 512, 1024])
def test_istft_roundtrip_and_oracle(L, n_fft):
    hop = n_fft // 2
    B, T = 2, 5000
    g = torch.Generator().manual_seed(n_fft)
    F = 1 + T // hop
    p = torch.randn(B, n_fft + 2, F, generator=g)
    ref = O.convert_length(O.istft_packed(p, n_fft, hop), T)
    ld = n_fft + 4
    rows = torch.zeros(B * F, ld)
    rows[:, : n_fft + 2] = p.transpose(1, 2).reshape(B * F, n_fft + 2)
    fr = torch.empty(B * F, n_fft, device="cuda")
    L.irfft_frames(rows.cuda(), B * F, ld, n_fft, fr)
    out = torch.empty(B, T, device="cuda")
    L.ola_combine([fr], [n_fft], [hop], [F], None, None, out, B, T, False, 0.0, 0.0, False)
    assert rel_rms(out.cpu(), ref) < 3e-6
    # STFT -> iSTFT reconstructs the signal (size-independent property)
    x = audio_input(B, T, seed=1)
    pk = torch.empty(B * F, ld, device="cuda")
    L.stft(x.cuda(), B, T, T, n_fft, hop, L.SPEC_PACKED, pk, ld)
    L.irfft_frames(pk, B * F, ld, n_fft, fr)
    L.ola_combine([fr], [n_fft], [hop], [F], None, None, out, B, T, False, 0.0, 0.0, False)
    valid = hop * (F - 1)
    assert rel_rms(out.cpu()[:, :valid], x[:, :valid]) < 3e-6
    assert float(out[:, valid:].abs().max()) == 0.0


def test_ola_combine_mean_euler_clamp(L):
    B, T = 2, 4096
    g = torch.Generator().manual_seed(5)
    cfgs = [(512, 256), (256, 128), (128, 64)]
    frs, outs = [], []
    for n, h in cfgs:
        F = 1 + T // h
        p = torch.randn(B, n + 2, F, generator=g) * 3
        outs.append(O.convert_length(O.istft_packed(p, n, h), T))
        rows = p.transpose(1, 2).reshape(B * F, n + 2).contiguous()
        fr = torch.empty(B * F, n, device="cuda")
        L.irfft_frames(rows.cuda(), B * F, n + 2, n, fr)
        frs.append(fr)
    x = torch.randn(B, T, generator=g)
    pred = torch.stack(outs, 1).mean(1)
    t, dt = 0.25, 0.25
    ref = (x + (pred - x) / (1 - torch.tensor(t)) * torch.tensor(dt)).clamp(-1, 1)
    xg = x.cuda()
    L.ola_combine(frs, [c[0] for c in cfgs], [c[1] for c in cfgs], [1 + T // c[1] for c in cfgs],
                  None, xg, xg, B, T, True, t, dt, True)
    assert rel_rms(xg.cpu(), ref) < 3e-6
    w = torch.rand(B, 3, generator=g)
    out = torch.empty(B, T, device="cuda")
    L.ola_combine(frs, [c[0] for c in cfgs], [c[1] for c in cfgs], [1 + T // c[1] for c in cfgs],
                  w.cuda(), None, out, B, T, False, 0.0, 0.0, False)
    assert rel_rms(out.cpu(), (torch.stack(outs, 1) * w[:, :, None]).sum(1)) < 3e-6


@pytest.mark.parametrize("C", [384, 512, 768])
def test_biasnorm_and_block_pre(L, C):
    B, T, Tc, factor = 2, 37, 9, 4
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, T, generator=g)
    bias = torch.randn(C, generator=g) * 0.1
    ls = torch.tensor(0.7)
    ref = O.bias_norm(x, bias, ls)
    xr = x.transpose(1, 2).reshape(B * T, C).contiguous().cuda()
    y = torch.empty_like(xr)
    L.biasnorm(xr, B * T, C, C, bias.cuda(), ls.cuda(), y, C)
    assert rel_rms(y.cpu().view(B, T, C).transpose(1, 2), ref) < 2e-6
    # fused prologue
    dw = torch.randn(C, 1, 7, generator=g) * 0.3
    dwb = torch.randn(C, generator=g) * 0.1
    cond = torch.randn(B, C, Tc, generator=g)
    zero_vec = torch.randn(C, generator=g)
    ts = torch.randn(B, C, generator=g) * 0.3
    lens = torch.tensor([T, T - 11])
    mask = (torch.arange(T)[None] < lens[:, None]).float()
    conv = torch.nn.functional.conv1d(x * mask[:, None], dw, dwb, padding=3, groups=C)
    z = O.bias_norm(conv, bias, ls)
    cup = torch.cat([cond.repeat_interleave(factor, 2),
                     zero_vec[None, :, None].expand(B, C, T - Tc * factor)], 2)
    ref = tf32_round(((z + cup) * (1 + ts[:, :, None])).contiguous())
    crow = torch.cat([cond.transpose(1, 2).reshape(B * Tc, C), zero_vec[None]], 0).contiguous().cuda()
    dwT = dw[:, 0, :].t().contiguous().cuda()
    out = torch.empty(B * T, C, device="cuda")
    conv_out = torch.empty(B * T, C, device="cuda")
    L.block_pre(xr, B, T, C, C, dwT, dwb.cuda(), bias.cuda(), ls.cuda(), mask.reshape(-1).cuda(), crow, C,
                Tc, factor, B * Tc, ts.cuda(), C, out, C, conv_out, None)
    assert rel_rms(out.cpu().view(B, T, C).transpose(1, 2), ref) < 3e-4   # tf32 rounding ties
    assert rel_rms(conv_out.cpu().view(B, T, C).transpose(1, 2), conv) < 2e-6
    # fp16 destination (operand of the kind::f16 block GEMMs): same values, RN to 11 significant bits
    out16 = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    L.block_pre(xr, B, T, C, C, dwT, dwb.cuda(), bias.cuda(), ls.cuda(), mask.reshape(-1).cuda(), crow, C,
                Tc, factor, B * Tc, ts.cuda(), C, out16, C, None, None)
    full = ((z + cup) * (1 + ts[:, :, None])).contiguous()
    assert rel_rms(out16.float().cpu().view(B, T, C).transpose(1, 2), full) < 3e-4
    assert float((out16.float().cpu() - out.cpu()).abs().max()) <= float(full.abs().max()) * 2 ** -10


def test_time_embedding_path(L):
    B, dim = 5, 512
    g = torch.Generator().manual_seed(0)
    t = torch.rand(B, generator=g)
    ref = O.sinusoidal_pos_emb(t, dim)
    half = dim // 2
    freqs = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    emb = torch.empty(B, dim, device="cuda")
    L.time_sinusoid(t.cuda(), B, dim, freqs.cuda(), 1000.0, emb)
    assert float((emb.cpu() - ref).abs().max()) < 2e-4        # sin/cos of arguments up to 1000
    W = torch.randn(700, dim, generator=g) / math.sqrt(dim)
    bb = torch.randn(700, generator=g)
    out = torch.empty(B, 700, device="cuda")
    L.linear_small(emb, B, dim, dim, W.cuda(), dim, bb.cuda(), 700, L.ACT_SILU, out, 700)
    assert rel_rms(out.cpu(), torch.nn.functional.silu(emb.cpu() @ W.t() + bb)) < 2e-6


def test_im2col_and_masks(L):
    B, C, T = 2, 100, 13
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, T, generator=g)
    ld = 304
    out = torch.empty(B * T, ld, device="cuda")
    L.im2col_cf(x.cuda(), B, C, T, 3, out, ld, 0)
    xp = torch.nn.functional.pad(x, (1, 1))
    ref = torch.zeros(B, T, ld)
    for k in range(3):
        ref[:, :, k * C:(k + 1) * C] = xp[:, :, k:k + T].transpose(1, 2)
    assert torch.equal(out.cpu().view(B, T, ld), ref)
    lens = torch.tensor([1000, 777], dtype=torch.int32)
    m = torch.empty(2 * 9, device="cuda")
    L.frame_mask(lens.cuda(), 2, 9, 128, m)
    ref = (torch.arange(9)[None] < (1 + lens // 128)[:, None]).float()
    assert torch.equal(m.cpu().view(2, 9), ref)
    a = audio_input(3, 3000, seed=3) + 0.1
    pre = torch.empty(3, 2, device="cuda")
    L.dc_peak(a.cuda(), 3, 3000, 3000, pre)
    mean = a.mean(-1)
    sc = 0.8 / ((a - mean[:, None]).abs().max(-1)[0] + 1e-9)
    assert rel_rms(pre.cpu(), torch.stack([mean, sc], 1)) < 1e-5
