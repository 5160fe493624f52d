"""Generates tests/golden/plan_math.json from the REFERENCE's own src/common sources
(oracle/_ref/libfinufft_ref_common.so, built by oracle/build.py from /root/reference).
Run in the build container:  python tests/golden/make_golden.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build as ob  # noqa: E402

path = ob.build_ref()
assert path, "needs /root/reference"
R = C.CDLL(path)
R.ref_lowest_sigma.restype = C.c_double
R.ref_lowest_sigma.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double]
out = {"source": "flatironinstitute/finufft @ 9810998d src/common/{kernel,pswf,utils}.cpp",
       "kernel": [], "pswf": [], "polyfit": [], "lowest_sigma": [], "nhg_type3": []}
for tol, dim, typ, sig, isf in [(1e-6, 3, 1, 2.0, 1), (1e-5, 2, 2, 2.0, 1), (1e-9, 2, 1, 2.0, 0),
                                (1e-9, 1, 1, 2.0, 0), (1e-6, 3, 3, 2.0, 1), (1e-6, 3, 3, 1.25, 1),
                                (1e-3, 2, 1, 2.0, 1), (1e-12, 3, 1, 2.0, 0), (1e-4, 1, 2, 1.25, 0)]:
    ns, beta = C.c_int(), C.c_double()
    R.ref_kernel_ns_beta(C.c_double(tol), dim, typ, C.c_double(sig), isf, C.byref(ns), C.byref(beta))
    out["kernel"].append(dict(tol=tol, dim=dim, type=typ, sigma=sig, is_float=isf, ns=ns.value,
                              beta=beta.value))
x = np.linspace(-1, 1, 41)
for c in (4.6624, 14.087166941154068, 16.443361431346414, 23.511944901923446, 37.0):
    psi = np.zeros_like(x)
    R.ref_pswf(C.c_double(c), C.c_int64(x.size), x.ctypes.data_as(C.c_void_p),
               psi.ctypes.data_as(C.c_void_p))
    out["pswf"].append(dict(c=c, x=x.tolist(), psi=psi.tolist()))
for suf, dt in (("f32", np.float32), ("f64", np.float64)):
    for ns_, beta_ in ((6, 14.087166941154068), (7, 16.443361431346414), (10, 23.511944901923446)):
        n = min(19, ns_ + 3)
        for panel in (0, ns_ // 2, ns_ - 1):
            b = np.zeros(n, dtype=dt)
            getattr(R, "ref_polyfit_pswf_" + suf)(ns_, C.c_double(beta_), panel, n,
                                                  b.ctypes.data_as(C.c_void_p))
            out["polyfit"].append(dict(dtype=suf, ns=ns_, beta=beta_, panel=panel, n=n,
                                       coef=[float(v) for v in b]))
for args in [(1e-6, 3, 7, 1.1920929e-07, 512.0), (1e-5, 2, 6, 1.1920929e-07, 4096.0),
             (1e-9, 1, 10, 2.220446049250313e-16, 2e6), (1e-4, 2, 5, 1.1920929e-07, 100.0),
             (3e-6, 1, 7, 1.1920929e-07, 64.0)]:
    out["lowest_sigma"].append(dict(args=list(args), value=R.ref_lowest_sigma(*args)))
for args in [(2.0, 3.14159, 107.5, 7), (1.25, 1.0, 50.0, 8), (2.0, 0.0, 10.0, 5), (2.0, 2.0, 0.0, 4)]:
    nf, h, g = C.c_int64(), C.c_double(), C.c_double()
    R.ref_nhg_type3(C.c_double(args[0]), C.c_double(args[1]), C.c_double(args[2]), args[3],
                    C.byref(nf), C.byref(h), C.byref(g))
    out["nhg_type3"].append(dict(args=list(args), nf=nf.value, h=h.value, gam=g.value))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plan_math.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote plan_math.json")
