"""One-call NUFFTs on device arrays, mirroring python/cufinufft/cufinufft/_simple.py:4-95:
nufft{1,2,3}d{1,2,3}(points..., data, ...) -> plan, setpts, execute, destroy."""
from .plan import Plan


def _invoke(nufft_type, dim, pts, data, n_modes, out, eps, isign, freqs=None, **kw):
    if nufft_type == 1:
        if out is not None:
            n_modes = tuple(out.shape[-dim:])
        if n_modes is None:
            raise ValueError("n_modes or out must be given for type 1")
        n_trans = data.shape[0] if data.dim() == 2 else 1
        arg = n_modes
    elif nufft_type == 2:
        n_trans = data.shape[0] if data.dim() == dim + 1 else 1
        arg = tuple(data.shape[-dim:])
    else:
        n_trans = data.shape[0] if data.dim() == 2 else 1
        arg = dim
    dtype = "complex64" if str(data.dtype).endswith("complex64") else "complex128"
    plan = Plan(nufft_type, arg, n_trans, eps, isign, dtype, **kw)
    try:
        plan.setpts(*pts, *(freqs or ()))
        return plan.execute(data, out)
    finally:
        plan.destroy()


def nufft1d1(x, data, n_modes=None, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(1, 1, (x,), data, n_modes if n_modes is None or not isinstance(n_modes, int)
                   else (n_modes,), out, eps, isign, **kw)


def nufft1d2(x, data, out=None, eps=1e-6, isign=-1, **kw):
    return _invoke(2, 1, (x,), data, None, out, eps, isign, **kw)


def nufft2d1(x, y, data, n_modes=None, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(1, 2, (x, y), data, n_modes, out, eps, isign, **kw)


def nufft2d2(x, y, data, out=None, eps=1e-6, isign=-1, **kw):
    return _invoke(2, 2, (x, y), data, None, out, eps, isign, **kw)


def nufft3d1(x, y, z, data, n_modes=None, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(1, 3, (x, y, z), data, n_modes, out, eps, isign, **kw)


def nufft3d2(x, y, z, data, out=None, eps=1e-6, isign=-1, **kw):
    return _invoke(2, 3, (x, y, z), data, None, out, eps, isign, **kw)


def nufft1d3(x, data, s, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(3, 1, (x, None, None), data, None, out, eps, isign, freqs=(s,), **kw)


def nufft2d3(x, y, data, s, t, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(3, 2, (x, y, None), data, None, out, eps, isign, freqs=(s, t), **kw)


def nufft3d3(x, y, z, data, s, t, u, out=None, eps=1e-6, isign=1, **kw):
    return _invoke(3, 3, (x, y, z), data, None, out, eps, isign, freqs=(s, t, u), **kw)
