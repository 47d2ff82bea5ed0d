"""CPU emulation of the C-ABI calls flow2gan_b200/convwin.py makes (numpy restatement of the
documented operand addressing in include/flow2gan_b200.h), to validate the windowed-conv GEOMETRY
(padding, garbage rows, phase decomposition) against torch.nn.functional.conv2d without a GPU.
Development tool only -- the product path has no CPU fallback."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from flow2gan_b200 import _lib as L          # noqa: E402
from flow2gan_b200 import convwin            # noqa: E402


def arr(ptr, n):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,))


def pad2d(x_ptr, Nb, H, W, Cc, pn, ph_, pw_, Hl, Wp, ph, pw, slack, out, round_tf32=1):
    o = out.numpy()
    o[:] = 0
    ov = o[:Nb * Hl * Wp * Cc].reshape(Nb, Hl, Wp, Cc)
    for n in range(Nb):
        for h in range(H):
            for w in range(W):
                ov[n, h + ph, w + pw, :] = arr(x_ptr + 4 * (n * pn + h * ph_ + w * pw_), Cc)


def conv_w_pack(src, Co, Ci, taps, Co_pad, ld, dst, direction):
    assert direction == 0
    d = dst.numpy()
    d[:] = 0
    d[:Co, :taps * Ci] = src.reshape(Co, Ci, taps).permute(0, 2, 1).reshape(Co, taps * Ci).numpy()


def pack2d(src, rs, cs, rows, cols, dst, ld, ld_fill, rnd):
    assert src == dst


def act_bwd(dh, ld_dh, z, ld_z, slope, leaky, act, rows, cols, dz, ld_dz, g_bias, g_slope, round_tf32=0):
    d = dh.numpy()
    if act == L.ACT_LEAKY:
        zz = z.numpy()
        d[:, :cols] *= np.where(zz[:, :cols] > 0, 1.0, leaky).astype(np.float32)
    g_bias.numpy()[:cols] += d[:, :cols].sum(0)


def act_bwd_win(dy, Nb, Hl, R, Ho, Wo, z, ld_z, leaky, act, cols, cols_total, dz_ptr, ld_dz, guard_rows, g_bias,
                round_tf32=1):
    M = Nb * Hl * R
    full = arr(dz_ptr - 4 * guard_rows * ld_dz, (guard_rows + M) * ld_dz)
    full[:] = 0
    dzv = full[guard_rows * ld_dz:].reshape(Nb, Hl, R, ld_dz)
    o = dy.numpy().copy()
    if act == L.ACT_LEAKY:
        zz = z.numpy().reshape(Nb, Hl, R, ld_z)[:, :Ho, :Wo, :cols]
        o *= np.where(zz > 0, 1.0, leaky).astype(np.float32)
    dzv[:, :Ho, :Wo, :cols] = o
    g_bias.numpy()[:cols] += o.sum((0, 1, 2))


def conv_w_pack_dgrad(weight, Co, Ci, kh, kw, sw, Cop, out):
    g = convwin.WinGeom(0, 0, 0, Ci, Co, kh, kw, sw, 0, 0, 0, 0, 0, 0, 0, Cop, 0, 0, 0, 0, 0)
    o, off = out.numpy(), 0
    for phase in range(sw):
        wd = convwin.pack_dgrad_weight(weight, g, phase).contiguous().numpy().reshape(-1)
        o[off:off + wd.size] = wd
        off += wd.size
    assert off == o.size


def gemm_group(descs):
    for d in descs:
        M, N, K = d.M, d.N, d.K
        seg, shift, rows = d.a_seg_len, d.a_seg_shift, d.a_rows
        assert seg and seg % 32 == 0
        span = (rows - 1) * d.lda + seg
        a = arr(d.a, span)
        A = np.zeros((M, K), np.float32)
        if not d.a_mn:
            for k0 in range(0, K, seg):
                s = k0 // seg
                w = min(seg, K - k0)
                for m in range(M):
                    r = m + s * shift
                    if 0 <= r < rows:
                        A[m, k0:k0 + w] = a[r * d.lda: r * d.lda + w]
        else:
            for m0 in range(0, M, seg):
                s = m0 // seg
                w = min(seg, M - m0)
                for k in range(K):
                    r = k + s * shift
                    if 0 <= r < rows:
                        A[m0:m0 + w, k] = a[r * d.lda: r * d.lda + w]
        if d.b_mn:
            Bm = np.stack([arr(d.b + 4 * k * d.ldb, N) for k in range(K)], 0)        # (K, N)
        else:
            Bm = np.stack([arr(d.b + 4 * n * d.ldb, K) for n in range(N)], 1)        # (K, N)
        Cm = A.astype(np.float64) @ Bm.astype(np.float64)
        if d.bias:
            Cm += arr(d.bias, N)[None, :]
        if d.act == L.ACT_LEAKY:
            Cm = np.where(Cm > 0, Cm, d.leaky * Cm)
        for m in range(M):
            arr(d.c + 4 * m * d.ldc, N)[:] = Cm[m].astype(np.float32)


L.pad2d, L.conv_w_pack, L.pack2d, L.act_bwd, L.gemm_group = pad2d, conv_w_pack, pack2d, act_bwd, gemm_group
L.act_bwd_win, L.conv_w_pack_dgrad = act_bwd_win, conv_w_pack_dgrad
L.ptr = lambda t: None if t is None else t.data_ptr()


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check(Nb, H, W, C, Co, kh, kw, sw, ph, pw, leaky):
    torch.manual_seed(Nb + H + W + C)
    x = torch.randn(Nb, H, W + 5, C)
    wt = torch.randn(Co, C, kh, kw) / (C * kh * kw) ** 0.5
    b = torch.randn(Co)
    xg, wg, bg = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    y = convwin.conv2d_win(xg[:, :, 2:2 + W, :], wg, bg, sw, ph, pw, leaky)
    xr, wr, br = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    yr = torch.nn.functional.conv2d(xr[:, :, 2:2 + W, :].permute(0, 3, 1, 2), wr, br, (1, sw), (ph, pw))
    if leaky is not None:
        yr = torch.nn.functional.leaky_relu(yr, leaky)
    yr = yr.permute(0, 2, 3, 1)
    assert y.shape == yr.shape, (y.shape, yr.shape)
    g = torch.randn(yr.shape)
    (y * g).sum().backward()
    (yr * g).sum().backward()
    e = (rel(y.detach(), yr.detach()), rel(xg.grad, xr.grad), rel(wg.grad, wr.grad), rel(bg.grad, br.grad))
    print((Nb, H, W, C, Co, kh, kw, sw, ph, pw, leaky), "fwd %.1e dx %.1e dw %.1e db %.1e" % e)
    assert max(e) < 1e-5


if __name__ == "__main__":
    check(2, 5, 13, 32, 32, 3, 9, 2, 1, 4, 0.1)
    check(2, 5, 12, 32, 32, 3, 9, 2, 1, 4, 0.1)
    check(2, 4, 6, 32, 1, 3, 3, 1, 1, 1, None)
    check(2, 1, 20, 32, 48, 1, 5, 3, 0, 2, 0.1)
    check(2, 1, 7, 64, 40, 1, 5, 1, 0, 2, 0.1)
    check(1, 3, 4, 32, 33, 3, 9, 2, 1, 4, 0.1)
    check(2, 6, 3, 32, 8, 3, 1, 1, 1, 0, None)
    print("ok")
