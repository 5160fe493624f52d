"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy/ctypes front-end of oracle/liboracle.so (our CPU restatement of the reference
FINUFFT hot path) and, when present, oracle/_ref/libfinufft_ref_common.so (the reference's
own src/common sources).  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by finufft_b200/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_i64 = C.c_int64
_p = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _load():
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        from . import build as _b  # noqa
        _b.build_oracle()
    lib = C.CDLL(path)
    lib.orc_next235.restype = _i64
    lib.orc_next235.argtypes = [_i64, _i64]
    lib.orc_lowest_sigma.restype = C.c_double
    lib.orc_lowest_sigma.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.orc_eval_kernel_f32.restype = C.c_float
    lib.orc_eval_kernel_f32.argtypes = [C.c_float, C.c_int, C.c_int, _p]
    lib.orc_eval_kernel_f64.restype = C.c_double
    lib.orc_eval_kernel_f64.argtypes = [C.c_double, C.c_int, C.c_int, _p]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def ref_lib():
    """The reference's own src/common compiled into oracle/_ref, or None."""
    path = os.path.join(HERE, "_ref", "libfinufft_ref_common.so")
    if not os.path.exists(path):
        return None
    r = C.CDLL(path)
    r.ref_next235.restype = _i64
    r.ref_next235.argtypes = [_i64, _i64]
    r.ref_lowest_sigma.restype = C.c_double
    r.ref_lowest_sigma.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double]
    return r


_ref_dirft = False


def ref_dirft_lib():
    """The reference's own test/utils/dirft*.hpp + norms.hpp compiled into oracle/_ref, or None."""
    global _ref_dirft
    if _ref_dirft is False:
        path = os.path.join(HERE, "_ref", "libfinufft_ref_dirft.so")
        if not os.path.exists(path):
            try:
                from . import build as _b
                _b.build_ref_dirft()
            except Exception:
                pass
        if os.path.exists(path):
            r = C.CDLL(path)
            r.ref_relerrtwonorm.restype = C.c_double
            r.ref_relerrtwonorm.argtypes = [_i64, _p, _p]
            _ref_dirft = r
        else:
            _ref_dirft = None
    return _ref_dirft


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype in (np.float32, np.complex64):
        return "f32", np.float32, np.complex64
    if dtype in (np.float64, np.complex128):
        return "f64", np.float64, np.complex128
    raise TypeError(dtype)


def max_threads():
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------- plan-time maths
def next235(n, fac=1):
    return int(lib().orc_next235(int(n), int(fac)))


def gaussquad(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gaussquad(C.c_int(n), _ptr(x), _ptr(w))
    return x, w


def pswf(c, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    err = lib().orc_pswf(C.c_double(c), _i64(x.size), _ptr(x), _ptr(out))
    if err:
        raise RuntimeError(f"pswf error {err}")
    return out


def kernel_setup(tol, dim, type_, sigma, dtype, allow_small=False):
    s, _, _ = _suf(dtype)
    ns = C.c_int()
    beta = C.c_double()
    tolu = C.c_double()
    err = getattr(lib(), f"orc_kernel_setup_{s}")(
        C.c_double(tol), dim, type_, C.c_double(sigma), int(allow_small), C.byref(ns),
        C.byref(beta), C.byref(tolu))
    return err, ns.value, beta.value, tolu.value


def horner(ns, beta, tol, dtype):
    s, rt, _ = _suf(dtype)
    coef = np.zeros(19 * ns, dtype=rt)
    nc = C.c_int()
    err = getattr(lib(), f"orc_horner_{s}")(ns, C.c_double(beta), C.c_double(tol), _ptr(coef),
                                            C.byref(nc))
    if err:
        raise RuntimeError(f"horner error {err}")
    return coef[: nc.value * ns].reshape(nc.value, ns).copy(), nc.value


def polyfit_pswf(ns, beta, panel, n, dtype):
    s, rt, _ = _suf(dtype)
    out = np.zeros(n, dtype=rt)
    getattr(lib(), f"orc_polyfit_pswf_{s}")(ns, C.c_double(beta), panel, n, _ptr(out))
    return out


def eval_stencil(x1, coef):
    nc, ns = coef.shape
    s, rt, _ = _suf(coef.dtype)
    ker = np.zeros(ns, dtype=rt)
    ct = C.c_float if rt == np.float32 else C.c_double
    getattr(lib(), f"orc_eval_stencil_{s}")(ct(x1), ns, nc, _ptr(np.ascontiguousarray(coef)),
                                            _ptr(ker))
    return ker


def eval_kernel(x, coef):
    nc, ns = coef.shape
    s, rt, _ = _suf(coef.dtype)
    ct = C.c_float if rt == np.float32 else C.c_double
    cc = np.ascontiguousarray(coef)
    return getattr(lib(), f"orc_eval_kernel_{s}")(ct(x), ns, nc, _ptr(cc))


def fseries(nf, coef):
    nc, ns = coef.shape
    s, rt, _ = _suf(coef.dtype)
    out = np.zeros(nf // 2 + 1, dtype=rt)
    getattr(lib(), f"orc_fseries_{s}")(_i64(nf), ns, nc, _ptr(np.ascontiguousarray(coef)),
                                       _ptr(out))
    return out


def lowest_sigma(tol, dim, ns, eps_mach, gridlen):
    return lib().orc_lowest_sigma(tol, dim, ns, eps_mach, gridlen)


def nhg_type3(sigma, X, S, ns):
    nf = _i64()
    h = C.c_double()
    g = C.c_double()
    lib().orc_nhg_type3(C.c_double(sigma), C.c_double(X), C.c_double(S), ns, C.byref(nf),
                        C.byref(h), C.byref(g))
    return nf.value, h.value, g.value


# ----------------------------------------------------------------------------- hot-path stages
def fold_rescale(x, N):
    s, rt, _ = _suf(x.dtype)
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    getattr(lib(), f"orc_fold_rescale_{s}")(_i64(x.size), _ptr(x), _i64(N), _ptr(out))
    return out


def _nf3(nf):
    nf = list(nf) + [1] * (3 - len(nf))
    return nf


def bin_sort(x, y, z, nf):
    """Returns (perm, bins): the reference's stable bin-sort permutation and bin ids."""
    s, rt, _ = _suf(x.dtype)
    M = x.size
    n1, n2, n3 = _nf3(nf)
    perm = np.zeros(M, dtype=np.int64)
    bins = np.zeros(M, dtype=np.int64)
    getattr(lib(), f"orc_bin_sort_{s}")(_i64(M), _ptr(x), _ptr(y), _ptr(z), _i64(n1), _i64(n2),
                                        _i64(n3), _ptr(perm), _ptr(bins))
    return perm, bins


def spread(nf, x, y, z, c, perm, coef, nthr=1):
    s, rt, ct = _suf(x.dtype)
    dim = len(nf)
    nfa = np.array(_nf3(nf), dtype=np.int64)
    nc, ns = coef.shape
    fw = np.zeros(int(np.prod(nfa)), dtype=ct)
    c = np.ascontiguousarray(c, dtype=ct)
    getattr(lib(), f"orc_spread_{s}")(dim, _ptr(nfa), _i64(x.size), _ptr(x), _ptr(y), _ptr(z),
                                      _ptr(c), _ptr(perm), ns, nc,
                                      _ptr(np.ascontiguousarray(coef)), _ptr(fw), nthr)
    return fw


def interp(nf, x, y, z, fw, perm, coef, nthr=1):
    s, rt, ct = _suf(x.dtype)
    dim = len(nf)
    nfa = np.array(_nf3(nf), dtype=np.int64)
    nc, ns = coef.shape
    c = np.zeros(x.size, dtype=ct)
    fw = np.ascontiguousarray(fw, dtype=ct)
    getattr(lib(), f"orc_interp_{s}")(dim, _ptr(nfa), _i64(x.size), _ptr(x), _ptr(y), _ptr(z),
                                      _ptr(c), _ptr(perm), ns, nc,
                                      _ptr(np.ascontiguousarray(coef)), _ptr(fw), nthr)
    return c


def deconvolve(direction, ms, nf, modeord, phihat, fk=None, fw=None):
    """direction 1: fw -> fk (returns fk); direction 2: fk -> padded fw (returns fw)."""
    dim = len(ms)
    rt = phihat[0].dtype
    s, rt, ct = _suf(rt)
    msa = np.array(list(ms) + [1] * (3 - dim), dtype=np.int64)
    nfa = np.array(_nf3(nf), dtype=np.int64)
    ph = [np.ascontiguousarray(p) for p in phihat] + [None] * (3 - dim)
    if direction == 1:
        fw = np.ascontiguousarray(fw, dtype=ct).copy()
        fk = np.zeros(int(np.prod(msa)), dtype=ct)
    else:
        fk = np.ascontiguousarray(fk, dtype=ct).copy()
        fw = np.full(int(np.prod(nfa)), np.nan, dtype=ct)
    getattr(lib(), f"orc_deconvolve_{s}")(direction, dim, _ptr(msa), _ptr(nfa), modeord,
                                          _ptr(ph[0]), _ptr(ph[1]), _ptr(ph[2]), _ptr(fk),
                                          _ptr(fw))
    return fk if direction == 1 else fw


def fft(nf, sign, data, nthr=1):
    s, rt, ct = _suf(data.dtype)
    dim = len(nf)
    nfa = np.array(_nf3(nf), dtype=np.int64)
    out = np.ascontiguousarray(data, dtype=ct).copy()
    getattr(lib(), f"orc_fft_{s}")(dim, _ptr(nfa), sign, _ptr(out), nthr)
    return out


# ----------------------------------------------------------------------------- guru plan
class Plan:
    """Oracle guru plan (reference semantics: makeplan / setpts / execute / destroy)."""

    def __init__(self, type_, n_modes, iflag, ntr, tol, dtype, sigma=2.0, modeord=0,
                 spread_only=False, allow_small=True, nthr=1, dim=None):
        self.s, self.rt, self.ct = _suf(dtype)
        self.type = type_
        self.dim = len(n_modes) if dim is None else dim
        self.ntr = ntr
        self.allow_small = int(allow_small)
        nm = np.array(list(n_modes) + [1] * (3 - len(n_modes)), dtype=np.int64)
        self.n_modes = [int(v) for v in nm[: self.dim]]
        h = C.c_void_p()
        err = getattr(lib(), f"orc_makeplan_{self.s}")(
            type_, self.dim, _ptr(nm), iflag, ntr, C.c_double(tol), C.c_double(sigma), modeord,
            int(spread_only), self.allow_small, nthr, C.byref(h))
        self.err = err
        self.h = h if err == 0 else None
        if err:
            raise RuntimeError(f"oracle makeplan error {err}")
        self._keep = None
        ns, nc, beta, tolu = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        nf = (C.c_int64 * 3)()
        getattr(lib(), f"orc_plan_info_{self.s}")(self.h, C.byref(ns), C.byref(nc),
                                                  C.byref(beta), nf, C.byref(tolu))
        self.ns, self.nc, self.beta, self.tol = ns.value, nc.value, beta.value, tolu.value
        self.nf = [int(nf[i]) for i in range(self.dim)]

    def tables(self):
        coef = np.zeros(self.nc * self.ns, dtype=self.rt)
        ph = [np.zeros(n // 2 + 1, dtype=self.rt) for n in self.nf] if self.type != 3 else []
        args = [_ptr(p) for p in ph] + [None] * (3 - len(ph))
        getattr(lib(), f"orc_plan_tables_{self.s}")(self.h, _ptr(coef), *args)
        return coef.reshape(self.nc, self.ns), ph

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None, force_sort=True):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.rt)
                for a in (x, y, z, s, t, u)]
        self._keep = arrs
        self.M = arrs[0].size
        self.nk = 0 if arrs[3] is None else arrs[3].size
        err = getattr(lib(), f"orc_setpts_{self.s}")(
            self.h, _i64(self.M), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _i64(self.nk),
            _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]), self.allow_small, int(force_sort))
        if err:
            raise RuntimeError(f"oracle setpts error {err}")
        if self.type == 3:  # fine grid known only now
            ns, nc, beta, tolu = C.c_int(), C.c_int(), C.c_double(), C.c_double()
            nf = (C.c_int64 * 3)()
            getattr(lib(), f"orc_plan_info_{self.s}")(self.h, C.byref(ns), C.byref(nc),
                                                      C.byref(beta), nf, C.byref(tolu))
            self.nf = [int(nf[i]) for i in range(self.dim)]

    def perm(self):
        p = np.zeros(self.M, dtype=np.int64)
        getattr(lib(), f"orc_plan_perm_{self.s}")(self.h, _ptr(p))
        return p

    def execute(self, data, adjoint=False):
        """type 1/3: data=c -> returns fk; type 2: data=fk -> returns c."""
        data = np.ascontiguousarray(data, dtype=self.ct)
        nm = int(np.prod(self.n_modes)) if self.type != 3 else self.nk
        forward_in_c = (self.type != 2) != adjoint
        if forward_in_c:
            c = data.reshape(-1).copy()
            fk = np.zeros(self.ntr * nm, dtype=self.ct)
        else:
            fk = data.reshape(-1).copy()
            c = np.zeros(self.ntr * self.M, dtype=self.ct)
        err = getattr(lib(), f"orc_execute_{self.s}")(self.h, _ptr(c), _ptr(fk), int(adjoint))
        if err:
            raise RuntimeError(f"oracle execute error {err}")
        return fk if forward_in_c else c

    def destroy(self):
        if self.h is not None:
            getattr(lib(), f"orc_destroy_{self.s}")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ----------------------------------------------------------------------------- the reference itself
class _RefOpts(C.Structure):  # include/finufft_opts.h:32-71, field for field
    _fields_ = [
        ("modeord", C.c_int), ("spreadinterponly", C.c_int), ("debug", C.c_int),
        ("spread_debug", C.c_int), ("showwarn", C.c_int), ("nthreads", C.c_int),
        ("fftw", C.c_int), ("spread_sort", C.c_int), ("spread_kerevalmeth", C.c_int),
        ("spread_kerpad", C.c_int), ("upsampfac", C.c_double), ("spread_thread", C.c_int),
        ("maxbatchsize", C.c_int), ("spread_nthr_atomic", C.c_int),
        ("spread_max_sp_size", C.c_int), ("spread_kerformula", C.c_int),
        ("allow_eps_too_small", C.c_int), ("fftw_lock_fun", C.c_void_p),
        ("fftw_unlock_fun", C.c_void_p), ("fftw_lock_data", C.c_void_p),
    ]


_ref_full = False


def ref_full_lib():
    """oracle/_ref/libfinufft_ref.so: the reference's own CPU library compiled where it lies
    (oracle/build.py::build_ref_library), or None."""
    global _ref_full
    if _ref_full is False:
        path = os.path.join(HERE, "_ref", "libfinufft_ref.so")
        if not os.path.exists(path):
            try:
                from . import build as _b
                _b.build_ref_library()
            except Exception:
                pass
        _ref_full = C.CDLL(path) if os.path.exists(path) else None
        if _ref_full is not None:
            for pre, real in (("", C.c_double), ("f", C.c_float)):
                mk = getattr(_ref_full, f"finufft{pre}_makeplan")
                mk.argtypes = [C.c_int, C.c_int, C.POINTER(_i64), C.c_int, C.c_int, real,
                               C.POINTER(_p), C.POINTER(_RefOpts)]
                sp = getattr(_ref_full, f"finufft{pre}_setpts")
                sp.argtypes = [_p, _i64, _p, _p, _p, _i64, _p, _p, _p]
                for nm in ("execute", "execute_adjoint"):
                    getattr(_ref_full, f"finufft{pre}_{nm}").argtypes = [_p, _p, _p]
                getattr(_ref_full, f"finufft{pre}_destroy").argtypes = [_p]
            for s in ("f32", "f64"):
                f = getattr(_ref_full, f"ref_plan_perm_{s}")
                f.restype = _i64
                f.argtypes = [_p, _p, C.POINTER(C.c_int)]
                f = getattr(_ref_full, f"ref_plan_phihat_{s}")
                f.restype = _i64
                f.argtypes = [_p, C.c_int, _p]
                getattr(_ref_full, f"ref_plan_info_{s}").argtypes = [
                    _p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                    C.POINTER(C.c_double), C.POINTER(_i64 * 3), C.POINTER(C.c_double)]
    return _ref_full


def have_reference():
    return ref_full_lib() is not None


class RefPlan:
    """Guru plan of the REFERENCE CPU library (finufft[f]_makeplan / setpts / execute /
    destroy, include/finufft/finufft_eitherprec.h:55-67), same Python surface as `Plan`."""

    def __init__(self, type_, n_modes, iflag, ntr, tol, dtype, sigma=2.0, modeord=0,
                 spread_only=False, allow_small=True, nthr=1, dim=None, spread_sort=1):
        L = ref_full_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libfinufft_ref.so is not available")
        self.L = L
        self.s, self.rt, self.ct = _suf(dtype)
        self.pre = "f" if self.s == "f32" else ""
        self.type, self.ntr = type_, ntr
        self.dim = len(n_modes) if dim is None else dim
        nm = (_i64 * 3)(*(list(n_modes) + [1] * (3 - len(n_modes))))
        self.n_modes = [int(nm[i]) for i in range(self.dim)]
        o = _RefOpts()
        getattr(L, f"finufft{self.pre}_default_opts")(C.byref(o))
        o.upsampfac = sigma
        o.modeord = modeord
        o.spreadinterponly = int(spread_only)
        o.allow_eps_too_small = int(allow_small)
        o.nthreads = nthr
        o.spread_sort = spread_sort
        o.showwarn = 0
        self.h = _p()
        real = C.c_float if self.s == "f32" else C.c_double
        err = getattr(L, f"finufft{self.pre}_makeplan")(type_, self.dim, nm, iflag, ntr,
                                                        real(tol), C.byref(self.h), C.byref(o))
        self.err = err
        if err > 1:  # 1 = warning: eps too small, clamped (allow_eps_too_small)
            self.h = None
            raise RuntimeError(f"reference makeplan error {err}")
        self._keep = None
        self._info()

    def _info(self):
        ns, nc = C.c_int(), C.c_int()
        beta, sigma, tol = C.c_double(), C.c_double(), C.c_double()
        nf = (_i64 * 3)()
        getattr(self.L, f"ref_plan_info_{self.s}")(self.h, C.byref(ns), C.byref(nc),
                                                  C.byref(beta), C.byref(sigma), C.byref(nf),
                                                  C.byref(tol))
        self.ns, self.nc, self.beta, self.sigma, self.tol = (ns.value, nc.value, beta.value,
                                                             sigma.value, tol.value)
        self.nf = [int(nf[i]) for i in range(self.dim)]

    def phihat(self, d):
        n = getattr(self.L, f"ref_plan_phihat_{self.s}")(self.h, d, None)
        out = np.zeros(n, dtype=self.rt)
        getattr(self.L, f"ref_plan_phihat_{self.s}")(self.h, d, _ptr(out))
        return out

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None, force_sort=True):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.rt)
                for a in (x, y, z, s, t, u)]
        self._keep = arrs  # the reference keeps the user's pointers (plan.hpp:130)
        self.M = arrs[0].size
        self.nk = 0 if arrs[3] is None else arrs[3].size
        err = getattr(self.L, f"finufft{self.pre}_setpts")(
            self.h, _i64(self.M), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _i64(self.nk),
            _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]))
        if err > 1:
            raise RuntimeError(f"reference setpts error {err}")
        self._info()

    def perm(self):
        did = C.c_int()
        p = np.zeros(self.M, dtype=np.int64)
        getattr(self.L, f"ref_plan_perm_{self.s}")(self.h, _ptr(p), C.byref(did))
        self.did_sort = bool(did.value)
        return p

    def execute(self, data, adjoint=False):
        data = np.ascontiguousarray(data, dtype=self.ct)
        nm = int(np.prod(self.n_modes)) if self.type != 3 else self.nk
        forward_in_c = (self.type != 2) != adjoint
        if forward_in_c:
            c = data.reshape(-1).copy()
            fk = np.zeros(self.ntr * nm, dtype=self.ct)
        else:
            fk = data.reshape(-1).copy()
            c = np.zeros(self.ntr * self.M, dtype=self.ct)
        fn = getattr(self.L, f"finufft{self.pre}_execute" + ("_adjoint" if adjoint else ""))
        err = fn(self.h, _ptr(c), _ptr(fk))
        if err > 1:
            raise RuntimeError(f"reference execute error {err}")
        return fk if forward_in_c else c

    def destroy(self):
        if self.h is not None:
            getattr(self.L, f"finufft{self.pre}_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ----------------------------------------------------------------------------- direct sums
def dirft(type_, x, y, z, data, iflag, n_modes=None, s=None, t=None, u=None, nthr=0, impl=None):
    """Double-precision direct sums, CMCL mode order.  impl="ref": the reference's own
    test/utils/dirft{1,2,3}d.hpp compiled into oracle/_ref (the default whenever that library is
    present); impl="port": our restatement of them (orc_dirft*)."""
    nthr = nthr or max_threads()
    rl = ref_dirft_lib() if impl in (None, "ref") else None
    if impl == "ref" and rl is None:
        raise RuntimeError("oracle/_ref/libfinufft_ref_dirft.so is not available")
    L = rl if rl is not None else lib()
    pre = "ref_" if rl is not None else "orc_"
    sign = 1 if iflag >= 0 else -1
    dim = 1 + (y is not None) + (z is not None)
    xs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z)]
    data = np.ascontiguousarray(data, dtype=np.complex128)
    M = xs[0].size
    if type_ == 3:
        st = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
              for a in (s, t, u)]
        nk = st[0].size
        out = np.zeros(nk, dtype=np.complex128)
        getattr(L, pre + "dirft3")(dim, _i64(M), _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(data), sign,
                         _i64(nk), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), _ptr(out), nthr)
        return out
    ms = np.array(list(n_modes) + [1] * (3 - dim), dtype=np.int64)
    if type_ == 1:
        out = np.zeros(int(np.prod(ms)), dtype=np.complex128)
        getattr(L, pre + "dirft1")(dim, _i64(M), _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(data), sign,
                         _ptr(ms), _ptr(out), nthr)
        return out
    out = np.zeros(M, dtype=np.complex128)
    getattr(L, pre + "dirft2")(dim, _i64(M), _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(out), sign,
                     _ptr(ms), _ptr(data), nthr)
    return out


def relerr(a, b):
    """||a-b||_2 / ||b||_2, b = the trusted array: the reference's own relerrtwonorm
    (test/utils/norms.hpp:17-37) when oracle/_ref holds it, else the same formula in numpy."""
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    rl = ref_dirft_lib()
    if rl is not None and a.size == b.size and a.size > 0:
        aa = np.ascontiguousarray(a, dtype=np.complex128)
        bb = np.ascontiguousarray(b, dtype=np.complex128)
        return float(rl.ref_relerrtwonorm(_i64(bb.size), _ptr(bb), _ptr(aa)))
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) /
                 np.linalg.norm(b.astype(np.complex128)))
