"""Host-only view of the library's plan mathematics (no GPU needed): what makeplan decides."""
import ctypes as C

import numpy as np

from . import _lib


def kernel(tol, dim, nufft_type, sigma, dtype, allow_small=True):
    """Returns (err, ns, beta, table[nc, ns]) as libfinufft_b200 would choose them."""
    lib = _lib.load()
    is_f = np.dtype(dtype) in (np.dtype("float32"), np.dtype("complex64"))
    rt = np.float32 if is_f else np.float64
    ns, nc, beta = C.c_int(), C.c_int(), C.c_double()
    coef = np.zeros(19 * 16, dtype=rt)
    err = lib.b200_host_kernel(tol, dim, nufft_type, sigma, int(is_f), int(allow_small),
                               C.byref(ns), C.byref(beta), C.byref(nc),
                               coef.ctypes.data_as(C.c_void_p))
    if err:
        return err, 0, 0.0, None
    return 0, ns.value, beta.value, coef[: nc.value * ns.value].reshape(nc.value, ns.value).copy()


def fine_grid(sigma, modes, ns):
    return int(_lib.load().b200_host_fine_grid(sigma, modes, ns))


def fseries(nf, table):
    nc, ns = table.shape
    t = np.ascontiguousarray(table)
    out = np.zeros(nf // 2 + 1, dtype=t.dtype)
    err = _lib.load().b200_host_fseries(nf, ns, nc, int(t.dtype == np.float32),
                                        t.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p))
    if err:
        raise RuntimeError(f"fseries error {err}")
    return out
