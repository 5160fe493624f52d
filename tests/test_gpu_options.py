"""Option semantics and one-shot wrappers of the reference API, computed on the GPU through the
C ABI and compared with the reference CPU library / the oracle:

* gpu_sort = 0 (include/cufinufft_opts.h:11) and spread_sort = 0 (include/finufft_opts.h:41):
  setpts keeps the identity permutation like the reference's indexSort
  (include/finufft/spreadinterp.hpp:186-191) and the point-driven kernels run; results equal the
  sorted path's to rounding.
* the one-shot wrappers cufinufft[f]{1,2,3}d{1,2,3}[many] (include/cufinufft.h:45-186) and
  finufft[f]...(include/finufft/finufft_eitherprec.h:72-151): plan + setpts + execute + destroy
  in one call, compared with the guru result of the checker.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import make_points

pytestmark = pytest.mark.gpu


def _rand_c(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


def _checker(oracle):
    return oracle.RefPlan if oracle.have_reference() else oracle.Plan


@pytest.mark.parametrize("prec,tol", [("f", 1e-5), ("d", 1e-10)])
@pytest.mark.parametrize("dim,modes", [(1, (300,)), (2, (40, 52)), (3, (20, 18, 24))])
@pytest.mark.parametrize("type_", [1, 2])
def test_no_sort_option_device_api(cuda, oracle, prec, tol, dim, modes, type_):
    import finufft_b200 as F
    rt, ct = (np.float32, np.complex64) if prec == "f" else (np.float64, np.complex128)
    rng = np.random.default_rng(17)
    M = 20_000
    pts = make_points(rng, dim, M, rt, "wide")[:dim]
    gp = F.Plan(type_, modes, 1, tol, 1, ct, upsampfac=2.0, gpu_sort=0)
    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    assert gp.sort_path() == 3
    assert np.array_equal(gp.sort_permutation(), np.arange(M, dtype=np.uint32))
    kw = dict(spread_sort=0) if oracle.have_reference() else {}
    op = _checker(oracle)(type_, list(modes[::-1]), 1, 1, tol, rt, sigma=2.0, nthr=4, **kw)
    op.setpts(*(pts[::-1] + [None] * (3 - dim)))
    if oracle.have_reference():  # the reference's own unsorted state: identity, didSort false
        assert np.array_equal(op.perm(), np.arange(M))
        assert op.did_sort is False
    data = _rand_c(rng, (M,) if type_ == 1 else modes, ct)
    got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
    assert oracle.relerr(got, op.execute(data)) <= 2 * tol
    # same plan, sorted: equal to rounding
    gs = F.Plan(type_, modes, 1, tol, 1, ct, upsampfac=2.0, gpu_sort=1)
    gs.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    assert gs.sort_path() != 3
    got_s = gs.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
    assert oracle.relerr(got, got_s) <= (2e-6 if prec == "f" else 1e-13)
    for p in (gp, gs, op):
        p.destroy()


def test_no_sort_option_host_api(cuda, oracle):
    import finufft_b200 as F
    rng = np.random.default_rng(3)
    modes, M, tol = (36, 30), 15_000, 1e-9
    pts = make_points(rng, 2, M, np.float64)[:2]
    c = _rand_c(rng, (2, M), np.complex128)
    out = {}
    for sort in (0, 1, 2):
        hp = F.HostPlan(1, modes, 2, tol, 1, "complex128", upsampfac=2.0, spread_sort=sort)
        hp.setpts(*pts)
        assert (hp.sort_path() == 3) == (sort == 0)
        out[sort] = hp.execute(c)
        hp.destroy()
    op = _checker(oracle)(1, list(modes[::-1]), 1, 2, tol, np.float64, nthr=4)
    op.setpts(pts[1], pts[0])
    want = op.execute(c)
    for sort in (0, 1, 2):
        assert oracle.relerr(out[sort], want) <= 2 * tol


# ----------------------------------------------------------------------------- one-shot wrappers
def _dev(cuda, a):
    return cuda.from_numpy(np.ascontiguousarray(a)).cuda()


def test_oneshot_cufinufftf3d1(cuda, oracle):
    """cufinufftf3d1(M, x, y, z, c, iflag, eps, ms, mt, mu, fk, opts) - cufinufft.h:136-139."""
    import finufft_b200 as F
    lib = F.load()
    rng = np.random.default_rng(5)
    M, tol, (m1, m2, m3) = 40_000, 1e-5, (22, 30, 26)
    x, y, z = make_points(rng, 3, M, np.float32)
    c = _rand_c(rng, (M,), np.complex64)
    dx, dy, dz, dc = (_dev(cuda, a) for a in (x, y, z, c))
    fk = cuda.zeros((m3, m2, m1), dtype=cuda.complex64, device="cuda")
    f = lib.cufinufftf3d1
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                  C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    assert f(M, dx.data_ptr(), dy.data_ptr(), dz.data_ptr(), dc.data_ptr(), 1, tol, m1, m2, m3,
             fk.data_ptr(), None) == 0
    op = _checker(oracle)(1, [m1, m2, m3], 1, 1, tol, np.float32, nthr=4)
    op.setpts(x, y, z)
    assert oracle.relerr(fk.cpu().numpy(), op.execute(c)) <= 2 * tol


def test_oneshot_cufinufft2d2many(cuda, oracle):
    """cufinufft2d2many(ntr, M, x, y, c, iflag, eps, ms, mt, fk, opts) - cufinufft.h:95-98."""
    import finufft_b200 as F
    lib = F.load()
    rng = np.random.default_rng(6)
    ntr, M, tol, (m1, m2) = 3, 30_000, 1e-10, (48, 40)
    x, y, _ = make_points(rng, 2, M, np.float64)
    fk = _rand_c(rng, (ntr, m2, m1), np.complex128)
    dx, dy, dfk = (_dev(cuda, a) for a in (x, y, fk))
    c = cuda.zeros((ntr, M), dtype=cuda.complex128, device="cuda")
    f = lib.cufinufft2d2many
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                  C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    assert f(ntr, M, dx.data_ptr(), dy.data_ptr(), c.data_ptr(), -1, tol, m1, m2, dfk.data_ptr(),
             None) == 0
    op = _checker(oracle)(2, [m1, m2], -1, ntr, tol, np.float64, nthr=4)
    op.setpts(x, y)
    assert oracle.relerr(c.cpu().numpy(), op.execute(fk)) <= 2 * tol


def test_oneshot_finufftf3d3_and_1d1(cuda, oracle):
    """Host-pointer one-shots: finufftf3d3(M, x, y, z, c, iflag, eps, N, s, t, u, fk, opts) and
    finufft1d1(M, x, c, iflag, eps, ms, fk, opts) - finufft_eitherprec.h:72-151."""
    import finufft_b200 as F
    lib = F.load()
    rng = np.random.default_rng(8)
    M, N, tol = 20_000, 15_000, 1e-5
    x, y, z = make_points(rng, 3, M, np.float32)
    s, t, u = [(20.0 * rng.uniform(-1, 1, N)).astype(np.float32) for _ in range(3)]
    c = _rand_c(rng, (M,), np.complex64)
    fk = np.zeros(N, dtype=np.complex64)
    f = lib.finufftf3d3
    f.restype = C.c_int
    vp = C.c_void_p
    f.argtypes = [C.c_int64, vp, vp, vp, vp, C.c_int, C.c_float, C.c_int64, vp, vp, vp, vp, vp]
    p = lambda a: a.ctypes.data_as(vp)  # noqa: E731
    assert f(M, p(x), p(y), p(z), p(c), 1, tol, N, p(s), p(t), p(u), p(fk), None) in (0, 1)
    op = _checker(oracle)(3, [1, 1, 1], 1, 1, tol, np.float32, nthr=4, dim=3)
    op.setpts(x, y, z, s, t, u)
    assert oracle.relerr(fk, op.execute(c)) <= 2 * tol
    # 1D type 1, double
    M, ms, tol = 30_000, 700, 1e-11
    x = make_points(rng, 1, M, np.float64)[0]
    c = _rand_c(rng, (M,), np.complex128)
    fk = np.zeros(ms, dtype=np.complex128)
    g = lib.finufft1d1
    g.restype = C.c_int
    g.argtypes = [C.c_int64, vp, vp, C.c_int, C.c_double, C.c_int64, vp, vp]
    assert g(M, p(x), p(c), 1, tol, ms, p(fk), None) == 0
    op = _checker(oracle)(1, [ms], 1, 1, tol, np.float64, nthr=4)
    op.setpts(x)
    assert oracle.relerr(fk, op.execute(c)) <= 2 * tol


@pytest.mark.parametrize("prec,tol,dim,modes,M", [
    ("d", 1e-9, 2, (512, 512), 2_000),          # sparse: the FFT dominates, sigma drops to ~1.2
    ("f", 1e-4, 3, (96, 96, 96), 500),
    ("d", 1e-6, 1, (20_000,), 300),
    ("d", 1e-9, 2, (64, 64), 200_000),          # dense: sigma stays 2
])
def test_auto_upsampfac_host_api(cuda, oracle, prec, tol, dim, modes, M):
    """finufft_opts.upsampfac = 0 on the host API: sigma is chosen at setpts (reference
    include/finufft/setpts.hpp:107-161, heuristics.hpp:82-128) from the number of points; the plan
    is rebuilt for it and the result matches the reference library run at that same sigma."""
    import finufft_b200 as F
    rt, ct = (np.float32, np.complex64) if prec == "f" else (np.float64, np.complex128)
    rng = np.random.default_rng(33)
    pts = make_points(rng, dim, M, rt)[:dim]
    for type_ in (1, 2):
        hp = F.HostPlan(type_, modes, 1, tol, 1, ct, allow_eps_too_small=1)   # upsampfac = 0
        assert hp.info()["sigma"] == 2.0                                       # until setpts
        hp.setpts(*pts)
        sigma = hp.info()["sigma"]
        if M < 10_000:
            assert 1.15 <= sigma < 2.0, sigma
        else:
            assert sigma == 2.0
        data = _rand_c(rng, (M,) if type_ == 1 else modes, ct)
        got = hp.execute(data)
        op = _checker(oracle)(type_, list(modes[::-1]), 1, 1, tol, rt, sigma=sigma, nthr=4)
        op.setpts(*(pts[::-1] + [None] * (3 - dim)))
        assert op.ns == hp.info()["ns"] and op.nf == hp.info()["nf"]
        want = op.execute(data)
        # At the smallest feasible sigma the aliasing error itself sits at the tolerance and
        # 1/phihat amplifies the last bits of the window fit at the edge modes, so two correct
        # implementations differ by ~2*tol (measured 2.03e-9 at tol 1e-9, sigma 1.2027).  Bars:
        # within 4*tol of the reference library, and as close to the direct sum as it is.
        assert oracle.relerr(got, want) <= 4 * tol, (type_, sigma)
        lp = pts[::-1] + [None] * (3 - dim)
        truth = oracle.dirft(type_, lp[0], lp[1], lp[2], data.reshape(-1), 1,
                             n_modes=list(modes[::-1]))
        e_gpu, e_ref = oracle.relerr(got, truth), oracle.relerr(want, truth)
        print(f"\n[auto sigma {sigma:.4f} type {type_}] gpu-vs-direct {e_gpu:.2e}  "
              f"reference-vs-direct {e_ref:.2e}")
        assert e_gpu <= max(2 * tol, 1.5 * e_ref), (type_, sigma, e_gpu, e_ref)
        hp.destroy()
        op.destroy()
