"""Synthetic bench inputs = the reference perftest's data (perftest/randunif.h:29-73): built
from tools/native/randunif.cpp into tools/native/librandunif.so (g++, no CUDA)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "randunif.cpp")
LIB = os.path.join(HERE, "native", "librandunif.so")

STREAMS = dict(X=0xA0761D6478BD642F, Y=0xE7037ED1A0B428DB, Z=0x8EBC6AF09C88C6E3,
               C=0x589965CC75374CC3, FK=0xEB44ACCAB455D165, S=0x9E3779B97F4A7C15,
               T=0xC2B2AE3D27D4EB4F, U=0x165667B19E3779F9)
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", SRC,
                               "-o", LIB])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        for name, real in (("b200_randunif_f32", C.c_float), ("b200_randunif_f64", C.c_double)):
            f = getattr(_lib, name)
            f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, real, real, C.c_int]
            f.restype = None
    return _lib


def fill(out, stream, scale=1.0, shift=0.0, nthreads=0, first=0):
    """out (contiguous float32/float64 numpy array, any shape) <- values [first, first+size) of
    the stream scale * (shift + U(-1,1))."""
    lib = _load()
    assert out.flags.c_contiguous
    f = lib.b200_randunif_f32 if out.dtype == np.float32 else lib.b200_randunif_f64
    f(out.ctypes.data_as(C.c_void_p), first, out.size, C.c_uint64(STREAMS[stream] if isinstance(stream, str) else stream),
      scale, shift, nthreads)
    return out


def points(dim, M, rt, dist="uniform", nf=None, first=0):
    """Coordinates like perftest (scale pi, shift 0); clustered = SURVEY.md 8(d): iid uniform in
    [0, 8 h_d), h_d = 2 pi / nf_d, the same streams with scale 4 h_d and shift 1.
    Returns [x, y, z][:dim] (library order, x first)."""
    out = []
    for d, s in zip(range(dim), "XYZ"):
        a = np.empty(M, dtype=rt)
        if dist == "cluster":
            h = 2 * np.pi / nf[d]
            fill(a, s, 4 * h, 1.0, first=first)
        else:
            fill(a, s, np.pi, 0.0, first=first)
        out.append(a)
    return out


def strengths(n, ct, stream="C", first=0):
    """n complex values, re and im uniform in (-1,1): one fill over 2n reals (perftest.cpp:172-190)."""
    rt = np.float32 if np.dtype(ct) == np.complex64 else np.float64
    a = np.empty(2 * n, dtype=rt)
    fill(a, stream, first=2 * first)
    return a.view(ct)
