"""Host-side mirror of the reference's Python plan interfaces, bound to the B200 library.

``Plan``      device arrays (torch CUDA tensors); mirrors python/cufinufft/cufinufft/_plan.py:39-405
              (same constructor arguments, setpts / execute methods, C-ordered data so the
              LAST array axis is the library's fastest (x) dimension, _plan.py:211-237).
``HostPlan``  numpy arrays through the host-pointer ABI; mirrors python/finufft's Plan.

PyTorch is used only to own device memory and streams; every computation happens in
libfinufft_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib

_ERRORS = {
    2: "fine grid too large", 3: "fine grid smaller than 2*nspread", 7: "upsampfac <= 1",
    9: "ntrans invalid", 10: "type invalid", 11: "allocation failed", 12: "dim invalid",
    14: "size does not fit 32-bit device indexing", 15: "CUDA failure", 16: "plan invalid",
    19: "insufficient shared memory", 20: "number of points invalid", 21: "invalid argument",
    24: "kerformula invalid", 25: "unknown exception", 26: "eps too small", 27: "PSWF setup",
}


class NufftError(RuntimeError):
    def __init__(self, code, where):
        super().__init__(f"finufft_b200 {where} failed with error {code}: "
                         f"{_ERRORS.get(code, 'unknown')}")
        self.code = code


def _check(code, where):
    if code != 0:
        raise NufftError(code, where)


def _dtypes(dtype):
    dtype = np.dtype(dtype)
    if dtype in (np.dtype("complex64"), np.dtype("float32")):
        return "f", np.float32, np.complex64
    if dtype in (np.dtype("complex128"), np.dtype("float64")):
        return "", np.float64, np.complex128
    raise TypeError("dtype must be complex64 or complex128")


class _PlanCommon:
    def info(self, inner=False):
        """What the plan decided; inner=True: the inner type-2 plan of a type-3 plan (after
        setpts)."""
        out = _lib.PlanInfo()
        get = self._lib.b200_get_inner_plan_info if inner else self._lib.b200_get_plan_info
        _check(get(self._plan, C.byref(out)), "plan_info")
        d = self.dim
        return dict(ns=out.ns, nc=out.nc, sigma=out.sigma, beta=out.beta, tol=out.tol,
                    nf=[out.nf[i] for i in range(d)], ms=[out.ms[i] for i in range(d)],
                    nbins=[out.nbins[i] for i in range(d)], M=out.M, nsub=out.nsub,
                    batch=out.batch, is_float=bool(out.is_float))

    def sort_permutation(self):
        """setpts result: sorted position -> user index (host numpy uint32)."""
        M = self.info()["M"]
        out = np.zeros(M, dtype=np.uint32)
        _check(self._lib.b200_get_sort_permutation(self._plan, out.ctypes.data_as(C.c_void_p)),
               "get_sort")
        return out

    def raw_sort_order(self):
        """The order the kernels work in (device array as setpts left it)."""
        M = self.info()["M"]
        out = np.zeros(M, dtype=np.uint32)
        _check(self._lib.b200_get_raw_sort_order(self._plan, out.ctypes.data_as(C.c_void_p)),
               "get_raw_sort")
        return out

    def sort_path(self):
        """0 counting sort, 1 partition sort, 2 stable radix sort (last setpts)."""
        v = C.c_int()
        _check(self._lib.b200_get_sort_path(self._plan, C.byref(v)), "get_sort_path")
        return v.value

    def window_table(self):
        i = self.info()
        out = np.zeros((i["nc"], i["ns"]), dtype=self._real)
        _check(self._lib.b200_get_window_table(self._plan, out.ctypes.data_as(C.c_void_p)),
               "get_table")
        return out

    def enable_profiling(self, on=True):
        _check(self._lib.b200_enable_profiling(self._plan, int(on)), "enable_profiling")

    def stage_ms(self):
        """dict of ms for the last execute / setpts (CUDA events on the plan's stream)."""
        ms = (C.c_float * 5)()
        _check(self._lib.b200_get_stage_ms(self._plan, C.byref(ms)), "stage_ms")
        return dict(spreadinterp=ms[0], fft=ms[1], deconv=ms[2], execute=ms[3], setpts=ms[4])

    def launch_count(self):
        n = C.c_uint64()
        _check(self._lib.b200_get_launch_count(self._plan, C.byref(n)), "launch_count")
        return int(n.value)

    def phihat(self, d):
        """Fourier series of the window for LIBRARY dimension d (0 = x = fastest)."""
        nf = self.info()["nf"][d]
        out = np.zeros(nf // 2 + 1, dtype=self._real)
        _check(self._lib.b200_get_phihat(self._plan, d, out.ctypes.data_as(C.c_void_p)),
               "get_phihat")
        return out


class Plan(_PlanCommon):
    """NUFFT plan on device arrays (torch CUDA tensors).

    Args mirror cufinufft.Plan: nufft_type (1, 2 or 3), n_modes (tuple, or the dimension as an
    int for type 3), n_trans, eps, isign (default +1 for types 1 and 3, -1 for type 2), dtype
    ('complex64' | 'complex128'), and keyword options named as the fields of cufinufft_opts
    (upsampfac, gpu_maxsubprobsize, gpu_spreadinterponly, gpu_maxbatchsize, gpu_device_id,
    gpu_stream, modeord, debug, ...).
    """

    def __init__(self, nufft_type, n_modes, n_trans=1, eps=1e-6, isign=None,
                 dtype="complex64", **kwargs):
        import torch  # device memory and streams only
        self._torch = torch
        self._lib = _lib.load()
        self._pre, self._real, self._cplx = _dtypes(dtype)
        if isign is None:
            isign = -1 if nufft_type == 2 else +1
        if isinstance(n_modes, int):
            if nufft_type == 3:
                self.dim, n_modes = n_modes, (1,) * n_modes
            else:
                n_modes = (n_modes,)
                self.dim = 1
        else:
            n_modes = tuple(int(n) for n in n_modes)
            self.dim = len(n_modes)
        self.type, self.n_modes, self.n_trans = nufft_type, n_modes, n_trans
        self.isign, self.eps = isign, eps
        opts = _lib.CufinufftOpts()
        self._lib.cufinufft_default_opts(C.byref(opts))
        for k, v in kwargs.items():
            if not hasattr(opts, k):
                raise TypeError(f"invalid option '{k}'")
            setattr(opts, k, v)
        self._opts = opts
        self.device = torch.device("cuda", opts.gpu_device_id)
        # C order: last python axis = library x (fastest)
        nm = (C.c_int64 * 3)(*(list(n_modes[::-1]) + [1] * (3 - self.dim)))
        self._plan = C.c_void_p()
        mk = getattr(self._lib, f"cufinufft{self._pre}_makeplan")
        real = C.c_float if self._pre == "f" else C.c_double
        _check(mk(nufft_type, self.dim, nm, isign, n_trans, real(eps), C.byref(self._plan),
                  C.byref(opts)), "makeplan")
        self._refs = []
        self.M = self.nk = 0

    # ------------------------------------------------------------------
    def _real_tensor(self, a, name):
        torch = self._torch
        if a is None:
            return None
        if not (torch.is_tensor(a) and a.is_cuda):
            raise TypeError(f"{name} must be a CUDA tensor")
        want = torch.float32 if self._pre == "f" else torch.float64
        if a.dtype != want:
            raise TypeError(f"{name} must have dtype {want}")
        return a.contiguous()

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None):
        pts = [self._real_tensor(a, n) for a, n in zip((x, y, z), "xyz")][: self.dim]
        if any(p is None for p in pts):
            raise TypeError(f"need {self.dim} coordinate arrays")
        M = pts[0].numel()
        if any(p.numel() != M for p in pts):
            raise TypeError("coordinate arrays must have equal length")
        frq = [self._real_tensor(a, n) for a, n in zip((s, t, u), "stu")][: self.dim]
        nk = 0
        if self.type == 3:
            if any(f is None for f in frq):
                raise TypeError(f"type 3 needs {self.dim} frequency arrays")
            nk = frq[0].numel()
        pts, frq = pts[::-1], frq[::-1]  # python axis order -> library x,y,z
        ptr = [p.data_ptr() for p in pts] + [None] * (3 - self.dim)
        fptr = ([f.data_ptr() for f in frq] if self.type == 3 else []) + [None] * 3
        sp = getattr(self._lib, f"cufinufft{self._pre}_setpts")
        _check(sp(self._plan, M, ptr[0], ptr[1], ptr[2], nk, fptr[0], fptr[1], fptr[2]), "setpts")
        self._refs = pts + frq  # the plan may read these again: keep them alive
        self.M, self.nk = M, nk

    def execute(self, data, out=None):
        torch = self._torch
        want = torch.complex64 if self._pre == "f" else torch.complex128
        if not (torch.is_tensor(data) and data.is_cuda and data.dtype == want):
            raise TypeError(f"data must be a CUDA tensor of dtype {want}")
        data = data.contiguous()
        batch = (self.n_trans,) if self.n_trans > 1 else ()
        if self.type == 2:
            in_shape, out_shape = batch + self.n_modes, batch + (self.M,)
        elif self.type == 1:
            in_shape, out_shape = batch + (self.M,), batch + self.n_modes
        else:
            in_shape, out_shape = batch + (self.M,), batch + (self.nk,)
        if tuple(data.shape) != in_shape:
            raise TypeError(f"data must have shape {in_shape}, got {tuple(data.shape)}")
        if out is None:
            out = torch.empty(out_shape, dtype=want, device=data.device)
        elif tuple(out.shape) != out_shape or out.dtype != want or not out.is_contiguous():
            raise TypeError(f"out must be contiguous {want} of shape {out_shape}")
        ex = getattr(self._lib, f"cufinufft{self._pre}_execute")
        if self.type == 2:
            _check(ex(self._plan, out.data_ptr(), data.data_ptr()), "execute")
        else:
            _check(ex(self._plan, data.data_ptr(), out.data_ptr()), "execute")
        return out

    def destroy(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            de = getattr(self._lib, f"cufinufft{self._pre}_destroy")
            de(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class HostPlan(_PlanCommon):
    """NUFFT plan on host numpy arrays through finufft[f]_* (mirrors finufft.Plan).

    Keyword options are fields of finufft_opts (upsampfac, modeord, spreadinterponly,
    maxbatchsize, allow_eps_too_small, debug, ...).
    """

    def __init__(self, nufft_type, n_modes_or_dim, n_trans=1, eps=1e-6, isign=None,
                 dtype="complex128", **kwargs):
        self._lib = _lib.load()
        self._pre, self._real, self._cplx = _dtypes(dtype)
        if isign is None:
            isign = -1 if nufft_type == 2 else +1
        if isinstance(n_modes_or_dim, int):
            if nufft_type == 3:
                self.dim, n_modes = n_modes_or_dim, (1,) * n_modes_or_dim
            else:
                self.dim, n_modes = 1, (n_modes_or_dim,)
        else:
            n_modes = tuple(int(n) for n in n_modes_or_dim)
            self.dim = len(n_modes)
        self.type, self.n_modes, self.n_trans = nufft_type, n_modes, n_trans
        opts = _lib.FinufftOpts()
        getattr(self._lib, f"finufft{self._pre}_default_opts")(C.byref(opts))
        for k, v in kwargs.items():
            if not hasattr(opts, k):
                raise TypeError(f"invalid option '{k}'")
            setattr(opts, k, v)
        nm = (C.c_int64 * 3)(*(list(n_modes[::-1]) + [1] * (3 - self.dim)))
        self._plan = C.c_void_p()
        real = C.c_float if self._pre == "f" else C.c_double
        mk = getattr(self._lib, f"finufft{self._pre}_makeplan")
        _check(mk(nufft_type, self.dim, nm, isign, n_trans, real(eps), C.byref(self._plan),
                  C.byref(opts)), "makeplan")
        self.M = self.nk = 0

    def setpts(self, x, y=None, z=None, s=None, t=None, u=None):
        pts = [None if a is None else np.ascontiguousarray(a, dtype=self._real)
               for a in (x, y, z)][: self.dim]
        frq = [None if a is None else np.ascontiguousarray(a, dtype=self._real)
               for a in (s, t, u)][: self.dim]
        M = pts[0].size
        nk = frq[0].size if self.type == 3 else 0
        pts, frq = pts[::-1], frq[::-1]
        p = [a.ctypes.data_as(C.c_void_p) for a in pts] + [None] * (3 - self.dim)
        f = ([a.ctypes.data_as(C.c_void_p) for a in frq] if self.type == 3 else []) + [None] * 3
        sp = getattr(self._lib, f"finufft{self._pre}_setpts")
        _check(sp(self._plan, M, p[0], p[1], p[2], nk, f[0], f[1], f[2]), "setpts")
        self._refs = pts + frq
        self.M, self.nk = M, nk

    def _run(self, data, out, adjoint):
        data = np.ascontiguousarray(data, dtype=self._cplx)
        batch = (self.n_trans,) if self.n_trans > 1 else ()
        modes = batch + (self.n_modes if self.type != 3 else (self.nk,))
        points = batch + (self.M,)
        c_is_input = (self.type != 2) != adjoint
        in_shape, out_shape = (points, modes) if c_is_input else (modes, points)
        if tuple(data.shape) != in_shape:
            raise TypeError(f"data must have shape {in_shape}, got {tuple(data.shape)}")
        if out is None:
            out = np.zeros(out_shape, dtype=self._cplx)
        ex = getattr(self._lib, f"finufft{self._pre}_execute" + ("_adjoint" if adjoint else ""))
        c, fk = (data, out) if c_is_input else (out, data)
        _check(ex(self._plan, c.ctypes.data_as(C.c_void_p), fk.ctypes.data_as(C.c_void_p)),
               "execute")
        return out

    def execute(self, data, out=None):
        return self._run(data, out, False)

    def execute_adjoint(self, data, out=None):
        return self._run(data, out, True)

    def destroy(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            getattr(self._lib, f"finufft{self._pre}_destroy")(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
