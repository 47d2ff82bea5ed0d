"""Host emulation of the SIMT kernels written against csrc/simt.cuh (datapath.cu, losses.cu: thread-
independent, run as sequential loops; optim.cu: cooperative -- shared memory, __syncthreads, warp
shuffles, atomics -- run with one host thread per warp whose 32 lanes are ucontext fibers): the SAME source is compiled
with g++ -DF2G_HOST_EMUL so its index arithmetic and rounding can be checked on a box without a GPU.  Test infrastructure only; the product library
never contains this build and flow2gan_b200/_lib.py never loads it."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flow2gan_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")

_lib = None
REVERSE = os.environ.get("F2G_EMUL_REVERSE", "0") == "1"      # run blocks / threads in descending order


class AvgTensor(C.Structure):
    _fields_ = [("avg", C.c_void_p), ("cur", C.c_void_p), ("numel", C.c_longlong),
                ("cur_is_f64", C.c_int), ("avg_is_f32", C.c_int)]


def available() -> bool:
    return shutil.which("g++") is not None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        cus = [os.path.join(CSRC, f) for f in ("datapath.cu", "losses.cu", "optim.cu", "blocks.cu", "spectral.cu", "train.cu",
                                                   "conv.cu")]
        srcs_extra = [os.path.join(CSRC, "fft_warp.cuh")]
        gemm = os.path.join(ROOT, "tests", "emul_gemm.cpp")      # host restatement of the tensor-core contraction
        srcs = cus + srcs_extra + [gemm, os.path.join(CSRC, "simt.cuh"), os.path.join(ROOT, "include", "flow2gan_b200.h")]
        h = hashlib.sha256()
        for s in srcs:
            h.update(open(s, "rb").read())
        os.makedirs(OUT, exist_ok=True)
        so = os.path.join(OUT, f"libf2g_emul_{h.hexdigest()[:12]}{'_rev' if REVERSE else ''}.so")
        if not os.path.exists(so):
            obj = so[:-3] + "_gemm.o"
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-c", gemm,
                                   "-o", obj])
            subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-ffp-contract=off", "-DF2G_HOST_EMUL",
                                   *(["-DF2G_EMUL_REVERSE"] if REVERSE else []), "-shared", "-fPIC", "-pthread", "-fopenmp", *cus,
                                   "-x", "none", obj, "-o", so])
        l = C.CDLL(so)
        from flow2gan_b200 import _lib as L          # the product's own signatures (include/flow2gan_b200.h)
        for name, (args, res) in L._SIGS.items():
            if hasattr(l, name):
                fn = getattr(l, name)
                fn.argtypes, fn.restype = args, res
        l.f2g_emul_last_error.restype = C.c_char_p
        _lib = l
    return _lib


def last_error() -> str:
    return lib().f2g_emul_last_error().decode()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pcm_decode(raw: bytes, fmt: int, channels: int, first: int, n: int, with_stats: bool = True):
    buf = np.frombuffer(bytearray(raw) + bytearray(8), dtype=np.uint8)     # aligned copy
    mono = np.full(n, np.nan, dtype=np.float32)
    stats = np.zeros(2, dtype=np.float32) if with_stats else None
    rc = lib().f2g_pcm_decode(_p(buf), fmt, channels, first, n, _p(mono), _p(stats), None)
    return rc, mono, stats


def gain_resample(x: np.ndarray, stats, norm_db: float, orig_r: int, new_r: int, width: int, taps: np.ndarray,
                  n_out: int):
    x = np.ascontiguousarray(x, dtype=np.float32)
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    out = np.full(n_out, np.nan, dtype=np.float32)
    rc = lib().f2g_gain_resample(_p(x), x.size, _p(stats), norm_db, orig_r, new_r, width, _p(taps), _p(out), n_out,
                                 None)
    return rc, out


def pcm16_encode(x: np.ndarray, clamp: bool):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(x.size, dtype=np.int16)
    rc = lib().f2g_pcm16_encode(_p(x), x.size, int(clamp), _p(out), None)
    return rc, out


def average_update(pairs, w_avg: float, w_cur: float, scale: float, chunk: int = 4096):
    """pairs: [(avg float64|float32 ndarray (updated in place), cur float32|float64 ndarray)]"""
    tab = (AvgTensor * len(pairs))()
    chunks = []
    for i, (a, c) in enumerate(pairs):
        assert a.dtype in (np.float64, np.float32) and a.flags.c_contiguous and c.flags.c_contiguous
        tab[i].avg, tab[i].cur, tab[i].numel = a.ctypes.data, c.ctypes.data, a.size
        tab[i].cur_is_f64, tab[i].avg_is_f32 = int(c.dtype == np.float64), int(a.dtype == np.float32)
        for ci in range((a.size + chunk - 1) // chunk):
            chunks += [i, ci]
    ch = np.asarray(chunks, dtype=np.int32)
    return lib().f2g_average_update(C.cast(tab, C.c_void_p), _p(ch), len(chunks) // 2, w_avg, w_cur, scale, None)


def native_fixture(monkeypatch):
    """Redirect flow2gan_b200._lib to the emulated library for every entry point it contains: the
    product's own ctypes wrappers (shape bookkeeping, descriptor structs) then run unchanged on CPU
    tensors.  Entry points that are not emulated (tcgen05 GEMMs, FFTs ...) raise AttributeError."""
    from flow2gan_b200 import _lib as L
    e = lib()
    monkeypatch.setattr(L, "lib", lambda: e)
    monkeypatch.setattr(L, "load", lambda: e)
    monkeypatch.setattr(L, "ptr", lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(L, "stream", lambda: None)
    monkeypatch.setattr(L, "require_cuda", lambda t, what: None)
    e.f2g_last_error = e.f2g_emul_last_error
    return L
