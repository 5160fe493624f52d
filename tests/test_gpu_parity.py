"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.

Bars (BASELINE.json north_star): bin indices / sort permutation BIT-EXACT; outputs within a
relative l2 error of 2 x the requested tolerance of the oracle (tolerances written below);
direct-sum checks with the reference's own thresholds (test/tolsweep.cpp:36,53-56); and, at
the full BASELINE sizes, size-independent properties (stable-sortedness of the permutation,
type-1/type-2 adjointness, linearity, spread mass conservation).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import make_points

pytestmark = pytest.mark.gpu


def _dt(prec):
    return (np.float32, np.complex64) if prec == "f" else (np.float64, np.complex128)


def _plans(F, O, type_, modes, ntr, tol, prec, modeord=0, sigma=2.0, **kw):
    rt, ct = _dt(prec)
    gp = F.Plan(type_, tuple(modes), ntr, tol, 1, ct, upsampfac=sigma, modeord=modeord, **kw)
    op = O.Plan(type_, list(modes[::-1]), 1, ntr, tol, rt, sigma=sigma, modeord=modeord,
                nthr=O.max_threads())
    return gp, op


def _rand_c(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


# ----------------------------------------------------------------------------- setpts
@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("dim,modes", [(1, (300,)), (2, (70, 52)), (3, (24, 30, 20))])
@pytest.mark.parametrize("kind", ["uniform", "cluster", "wide", "edges"])
def test_sort_permutation_bit_exact(cuda, oracle, prec, dim, modes, kind):
    import finufft_b200 as F
    rng = np.random.default_rng(11)
    rt, ct = _dt(prec)
    gp, op = _plans(F, oracle, 1, modes, 1, 1e-5 if prec == "f" else 1e-9, prec)
    M = 40_000
    pts = make_points(rng, dim, M, rt, kind, nf=gp.info()["nf"][::-1])[:dim]
    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    op.setpts(*(pts[::-1] + [None] * (3 - dim)))
    assert np.array_equal(gp.sort_permutation().astype(np.int64), op.perm())


@pytest.mark.parametrize("dim,modes", [(2, (70, 52)), (3, (24, 30, 20))])
def test_radix_sort_is_reference_permutation_on_device(cuda, oracle, dim, modes, monkeypatch):
    """B200_NUFFT_SORT=2 selects the stable LSD radix sort, whose raw device output (no per-bin
    restoring on the host) must already be the reference CPU permutation."""
    import finufft_b200 as F
    monkeypatch.setenv("B200_NUFFT_SORT", "2")
    monkeypatch.setenv("B200_NUFFT_SWEEP", "0")
    rng = np.random.default_rng(13)
    gp, op = _plans(F, oracle, 1, modes, 1, 1e-6, "f")
    for kind in ("uniform", "cluster"):
        pts = make_points(rng, dim, 50_000, np.float32, kind, nf=gp.info()["nf"][::-1])[:dim]
        gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
        op.setpts(*(pts[::-1] + [None] * (3 - dim)))
        assert np.array_equal(gp.sort_permutation().astype(np.int64), op.perm())


def test_sort_ragged_sizes(cuda, oracle):
    """M = 0, 1, 31, 33, 2047, 2049 ... (tile edges of the radix sort) and one-bin grids."""
    import finufft_b200 as F
    rng = np.random.default_rng(12)
    for modes in ((10, 12, 14), (7,), (8, 8)):
        dim = len(modes)
        gp, op = _plans(F, oracle, 1, modes, 1, 1e-4, "f")
        for M in (0, 1, 31, 33, 2047, 2048, 2049, 70_001):
            pts = make_points(rng, dim, M, np.float32)[:dim]
            gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
            op.setpts(*(pts[::-1] + [None] * (3 - dim)))
            assert np.array_equal(gp.sort_permutation().astype(np.int64), op.perm()), (modes, M)
            if M == 0:
                c = cuda.zeros(0, dtype=cuda.complex64, device="cuda")
                out = gp.execute(c)
                assert float(out.abs().max()) == 0.0  # reference test/dumbinputs.cpp:119-122


def test_plan_parameters_match_oracle(cuda, oracle):
    import finufft_b200 as F
    for type_, modes, tol, prec in [(1, (256, 256, 256), 1e-6, "f"), (2, (2048, 2048), 1e-5, "f"),
                                    (1, (512, 512), 1e-9, "d"), (1, (1000,), 1e-9, "d")]:
        gp, op = _plans(F, oracle, type_, modes, 1, tol, prec)
        i = gp.info()
        assert (i["ns"], i["nc"]) == (op.ns, op.nc) and i["nf"] == op.nf
        assert abs(i["beta"] - op.beta) < 1e-12
        ocoef, oph = op.tables()
        gcoef = gp.window_table()
        if prec == "f":
            assert np.array_equal(gcoef, ocoef)
        for d in range(len(modes)):
            ph = gp.phihat(d)
            assert np.max(np.abs(ph - oph[d])) <= 4 * np.finfo(ph.dtype).eps * abs(oph[d][0])
        gp.destroy()


# ----------------------------------------------------------------------------- transforms
CASES = [
    # type, modes, M, tol, prec
    (1, (1200,), 30_000, 1e-5, "f"), (2, (1200,), 30_000, 1e-5, "f"),
    (1, (80, 66), 60_000, 1e-5, "f"), (2, (80, 66), 60_000, 1e-5, "f"),
    (1, (40, 36, 30), 150_000, 1e-6, "f"), (2, (40, 36, 30), 150_000, 1e-6, "f"),
    (1, (1000,), 20_000, 1e-9, "d"), (2, (1000,), 20_000, 1e-9, "d"),
    (1, (64, 50), 40_000, 1e-9, "d"), (2, (64, 50), 40_000, 1e-9, "d"),
    (1, (22, 30, 26), 60_000, 1e-9, "d"), (2, (22, 30, 26), 60_000, 1e-9, "d"),
    (1, (30, 20, 24), 50_000, 1e-2, "f"), (2, (17, 19, 23), 50_000, 1e-3, "f"),  # narrow / odd
    (1, (16, 20, 18), 20_000, 1e-13, "d"), (2, (33, 29), 20_000, 1e-12, "d"),    # wide kernels
]


@pytest.mark.parametrize("type_,modes,M,tol,prec", CASES)
def test_transform_matches_oracle(cuda, oracle, type_, modes, M, tol, prec):
    """rel l2 error vs the oracle <= 2*tol (north_star bar)."""
    import finufft_b200 as F
    rng = np.random.default_rng(21)
    rt, ct = _dt(prec)
    dim = len(modes)
    gp, op = _plans(F, oracle, type_, modes, 1, tol, prec)
    pts = make_points(rng, dim, M, rt, "wide")[:dim]
    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    op.setpts(*(pts[::-1] + [None] * (3 - dim)))
    data = _rand_c(rng, (M,) if type_ == 1 else modes, ct)
    got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
    want = op.execute(data)
    assert oracle.relerr(got, want) <= 2 * tol
    gp.destroy()


@pytest.mark.parametrize("prec,tols", [("f", (1e-2, 1e-4, 1e-6)), ("d", (1e-3, 1e-6, 1e-9, 1e-12))])
@pytest.mark.parametrize("dim,modes", [(1, (50,)), (2, (25, 40)), (3, (10, 11, 12))])
@pytest.mark.parametrize("type_", [1, 2])
def test_transform_vs_direct_sum(cuda, oracle, prec, tols, dim, modes, type_):
    """reference test/tolsweep.cpp: M=500, these mode counts, pass iff
    relerr <= max(floor, slack*tol), floors float {2e-5,2e-5,1e-5}, double 3e-14, slack {4,4,5}."""
    import finufft_b200 as F
    floor = {"f": (2e-5, 2e-5, 1e-5), "d": (3e-14, 3e-14, 3e-14)}[prec][dim - 1]
    slack = (4, 4, 5)[dim - 1]
    rng = np.random.default_rng(31)
    rt, ct = _dt(prec)
    M = 500
    pts = make_points(rng, dim, M, rt)[:dim]
    for tol in tols:
        for isign in (+1, -1):
            gp = F.Plan(type_, modes, 1, tol, isign, ct, upsampfac=2.0)
            gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
            data = _rand_c(rng, (M,) if type_ == 1 else modes, ct)
            got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
            lp = pts[::-1] + [None] * (3 - dim)
            want = oracle.dirft(type_, lp[0], lp[1], lp[2], data.reshape(-1), isign,
                                n_modes=list(modes[::-1]))
            assert oracle.relerr(got, want) <= max(floor, slack * tol), (tol, isign)
            gp.destroy()


@pytest.mark.parametrize("modeord", [0, 1])
@pytest.mark.parametrize("type_", [1, 2])
def test_many_vectors_and_modeord(cuda, oracle, type_, modeord):
    """ntransf > maxbatch (several FFT batches), both mode orderings, even and odd sizes."""
    import finufft_b200 as F
    rng = np.random.default_rng(41)
    modes, M, ntr, tol = (21, 32), 30_000, 11, 1e-6
    gp, op = _plans(F, oracle, type_, modes, ntr, tol, "d", modeord=modeord, gpu_maxbatchsize=4)
    pts = make_points(rng, 2, M, np.float64)[:2]
    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    op.setpts(*(pts[::-1] + [None]))
    data = _rand_c(rng, (ntr, M) if type_ == 1 else (ntr,) + modes, np.complex128)
    got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
    want = op.execute(data)
    assert oracle.relerr(got, want) <= 2 * tol
    # stack equals single calls (reference test/finufft2dmany_test.cpp)
    gp1 = F.Plan(type_, modes, 1, tol, 1, np.complex128, upsampfac=2.0, modeord=modeord)
    gp1.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
    one = gp1.execute(cuda.from_numpy(data[3]).cuda()).cpu().numpy()
    assert oracle.relerr(got[3], one) < 1e-12


def test_spreadinterp_only(cuda, oracle):
    """gpu_spreadinterponly=1: spread straight into / interpolate straight from the user grid
    (reference src/fft.cpp:397-406, src/cuda/execute.cu:133-143); stage-level parity."""
    import finufft_b200 as F
    rng = np.random.default_rng(51)
    for prec, tol in (("f", 1e-5), ("d", 1e-9)):
        rt, ct = _dt(prec)
        for grid in ((90,), (48, 40), (30, 24, 36)):
            dim, M = len(grid), 20_000
            pts = make_points(rng, dim, M, rt)[:dim]
            lp = pts[::-1] + [None] * (3 - dim)
            err, ns, beta, tolu = oracle.kernel_setup(tol, dim, 1, 2.0, rt, True)
            coef, _ = oracle.horner(ns, beta, tolu, rt)
            perm, _ = oracle.bin_sort(lp[0], lp[1], lp[2], list(grid[::-1]))
            c = _rand_c(rng, (M,), ct)
            sp = F.Plan(1, grid, 1, tol, 1, ct, upsampfac=2.0, gpu_spreadinterponly=1)
            sp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
            fw = sp.execute(cuda.from_numpy(c).cuda()).cpu().numpy()
            want = oracle.spread(list(grid[::-1]), lp[0], lp[1], lp[2], c, perm, coef)
            assert oracle.relerr(fw, want) < (2e-6 if prec == "f" else 1e-14)
            g = _rand_c(rng, grid, ct)
            ip = F.Plan(2, grid, 1, tol, 1, ct, upsampfac=2.0, gpu_spreadinterponly=1)
            ip.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
            ci = ip.execute(cuda.from_numpy(g).cuda()).cpu().numpy()
            wanti = oracle.interp(list(grid[::-1]), lp[0], lp[1], lp[2], g.reshape(-1), perm, coef)
            assert oracle.relerr(ci, wanti) < (2e-6 if prec == "f" else 1e-14)


def _ns_of(oracle, tol, sigma, rt):
    return oracle.kernel_setup(tol, 2, 1, sigma, rt, True)[1]


# (precision, tol, sigma): together they hit every kernel width the 2D sweep kernels are built
# for (float 2..12, double 2..16), both window layouts (8 / 16 lanes) and both window steps
SWEEP2_WIDTHS = [("f", t, 2.0) for t in (3e-1, 3e-3, 1e-3, 3e-4, 3e-5, 3e-6, 2e-7)] + \
                [("f", t, 1.25) for t in (1e-3, 1e-4, 1e-5)] + \
                [("d", t, 2.0) for t in (1e-1, 3e-3, 1e-3, 3e-4, 1e-5, 3e-6, 1e-7, 1e-8, 1e-9,
                                         1e-10, 1e-11, 1e-12, 1e-13, 1e-14, 1e-15)] + \
                [("d", t, 1.25) for t in (1e-6, 1e-7, 1e-8, 1e-9)]


@pytest.mark.parametrize("kind", ["uniform", "cluster", "edges"])
def test_sweep2d_stage_parity_all_widths(cuda, oracle, kind):
    """2D row-sweep kernels (sweep2d.cuh), spread-only and interp-only, against the oracle's
    spread / interp on the same points, for every kernel width; plus the generic kernels
    (B200_NUFFT_SWEEP=0) as a second opinion."""
    import finufft_b200 as F
    rng = np.random.default_rng(55)
    grid, M = (96, 72), 30_000
    seen = set()
    for prec, tol, sigma in SWEEP2_WIDTHS:
        rt, ct = _dt(prec)
        err, ns, beta, tolu = oracle.kernel_setup(tol, 2, 1, sigma, rt, True)
        assert err == 0
        seen.add((prec, ns))
        coef, _ = oracle.horner(ns, beta, tolu, rt)
        pts = make_points(rng, 2, M, rt, kind, nf=grid)[:2]
        lp = pts[::-1] + [None]
        perm, _ = oracle.bin_sort(lp[0], lp[1], lp[2], list(grid[::-1]))
        c = _rand_c(rng, (M,), ct)
        g = _rand_c(rng, grid, ct)
        want_s = oracle.spread(list(grid[::-1]), lp[0], lp[1], lp[2], c, perm, coef)
        want_i = oracle.interp(list(grid[::-1]), lp[0], lp[1], lp[2], g.reshape(-1), perm, coef)
        # vs the oracle: float summation order; double: the two independent double-precision
        # fits of the window polynomials agree to ~1e-12 only (test_plan_parameters_match_oracle)
        bar = (3e-6 if kind == "uniform" else 3e-5) if prec == "f" else 2e-11
        dpts = [cuda.from_numpy(p).cuda() for p in pts]
        res = {}
        for sweep in ("1", "0"):
            os.environ["B200_NUFFT_SWEEP"] = sweep
            try:
                sp = F.Plan(1, grid, 1, tol, 1, ct, upsampfac=sigma, gpu_spreadinterponly=1)
                sp.setpts(*dpts)
                assert sp.info()["ns"] == ns
                fw = sp.execute(cuda.from_numpy(c).cuda()).cpu().numpy()
                ip = F.Plan(2, grid, 1, tol, 1, ct, upsampfac=sigma, gpu_spreadinterponly=1)
                ip.setpts(*dpts)
                ci = ip.execute(cuda.from_numpy(g).cuda()).cpu().numpy()
            finally:
                os.environ.pop("B200_NUFFT_SWEEP", None)
            assert oracle.relerr(fw, want_s) < bar, (prec, ns, sigma, sweep, "spread")
            assert oracle.relerr(ci, want_i) < bar, (prec, ns, sigma, sweep, "interp")
            res[sweep] = (fw, ci)
            sp.destroy()
            ip.destroy()
        # sweep vs generic kernels: same window values, only the summation order differs
        tight = bar if prec == "f" else 1e-13
        assert oracle.relerr(res["1"][0], res["0"][0]) < tight, (prec, ns, sigma, "spread")
        assert oracle.relerr(res["1"][1], res["0"][1]) < tight, (prec, ns, sigma, "interp")
    assert {n for p_, n in seen if p_ == "f"} >= set(range(2, 9))
    assert {n for p_, n in seen if p_ == "d"} >= set(range(3, 16))


@pytest.mark.parametrize("kind", ["uniform", "cluster", "edges"])
def test_sweep3d_stage_parity_all_widths(cuda, oracle, kind):
    """3D single-precision row-sweep kernels (sweep3d.cu), spread-only and interp-only, for every
    width they are built for (ns = 2..7) against the oracle's spread / interp on the same
    points and against the generic kernels (B200_NUFFT_SWEEP=0); wider kernels (sigma = 1.25
    at tight tolerance) must land on the generic kernels and still match."""
    import finufft_b200 as F
    rng = np.random.default_rng(57)
    grid, M = (28, 36, 48), 40_000
    seen = set()
    for tol, sigma in [(3e-1, 2.0), (3e-2, 2.0), (3e-3, 2.0), (3e-4, 2.0), (3e-5, 2.0),
                       (1e-6, 2.0), (1e-3, 1.25), (1e-5, 1.25), (1e-6, 1.25)]:
        rt, ct = np.float32, np.complex64
        err, ns, beta, tolu = oracle.kernel_setup(tol, 3, 1, sigma, rt, True)
        assert err == 0
        seen.add(ns)
        coef, _ = oracle.horner(ns, beta, tolu, rt)
        pts = make_points(rng, 3, M, rt, kind, nf=grid)
        lp = pts[::-1]
        perm, _ = oracle.bin_sort(lp[0], lp[1], lp[2], list(grid[::-1]))
        c = _rand_c(rng, (M,), ct)
        g = _rand_c(rng, grid, ct)
        want_s = oracle.spread(list(grid[::-1]), lp[0], lp[1], lp[2], c, perm, coef)
        want_i = oracle.interp(list(grid[::-1]), lp[0], lp[1], lp[2], g.reshape(-1), perm, coef)
        bar = 3e-6 if kind == "uniform" else 3e-5
        dpts = [cuda.from_numpy(p).cuda() for p in pts]
        res = {}
        for sweep in ("1", "0"):
            os.environ["B200_NUFFT_SWEEP"] = sweep
            try:
                sp = F.Plan(1, grid, 1, tol, 1, ct, upsampfac=sigma, gpu_spreadinterponly=1)
                sp.setpts(*dpts)
                assert sp.info()["ns"] == ns
                fw = sp.execute(cuda.from_numpy(c).cuda()).cpu().numpy()
                ip = F.Plan(2, grid, 1, tol, 1, ct, upsampfac=sigma, gpu_spreadinterponly=1)
                ip.setpts(*dpts)
                ci = ip.execute(cuda.from_numpy(g).cuda()).cpu().numpy()
            finally:
                os.environ.pop("B200_NUFFT_SWEEP", None)
            assert oracle.relerr(fw, want_s) < bar, (ns, sigma, sweep, "spread")
            assert oracle.relerr(ci, want_i) < bar, (ns, sigma, sweep, "interp")
            res[sweep] = (fw, ci)
            sp.destroy()
            ip.destroy()
        assert oracle.relerr(res["1"][0], res["0"][0]) < bar, (ns, sigma, "spread")
        assert oracle.relerr(res["1"][1], res["0"][1]) < bar, (ns, sigma, "interp")
    assert seen >= set(range(2, 8)), seen


def test_sweep2d_ragged_and_batched(cuda, oracle):
    """Few points (fewer than a chunk, one point, none), many vectors, odd grid sizes."""
    import finufft_b200 as F
    rng = np.random.default_rng(56)
    for prec, tol in (("f", 1e-5), ("d", 1e-10)):
        rt, ct = _dt(prec)
        for modes, M, ntr in (((37, 21), 1, 1), ((37, 21), 31, 3), ((16, 90), 33, 2),
                              ((50, 64), 7000, 4)):
            for type_ in (1, 2):
                gp, op = _plans(F, oracle, type_, modes, ntr, tol, prec)
                pts = make_points(rng, 2, M, rt, "wide")[:2]
                gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
                op.setpts(pts[1], pts[0], None)
                shape = (M,) if type_ == 1 else modes
                data = _rand_c(rng, ((ntr,) if ntr > 1 else ()) + shape, ct)
                got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
                want = op.execute(data)
                assert oracle.relerr(got, want) <= 2 * tol, (prec, modes, M, ntr, type_)
                gp.destroy()
        gp = F.Plan(1, (20, 20), 1, tol, 1, ct, upsampfac=2.0)
        gp.setpts(cuda.zeros(0, dtype=cuda.float32 if prec == "f" else cuda.float64, device="cuda"),
                  cuda.zeros(0, dtype=cuda.float32 if prec == "f" else cuda.float64, device="cuda"))
        out = gp.execute(cuda.zeros(0, dtype=cuda.complex64 if prec == "f" else cuda.complex128,
                                    device="cuda"))
        assert float(out.abs().max()) == 0.0
        gp.destroy()


@pytest.mark.parametrize("prec,tol", [("f", 1e-5), ("d", 1e-10)])
@pytest.mark.parametrize("dim,modes", [(1, (500,)), (2, (60, 44)), (3, (24, 20, 30))])
def test_staged_strength_permutation(cuda, oracle, prec, tol, dim, modes, monkeypatch):
    """Two-level permutation of the strengths (stage.cuh), forced on with 25 small windows:
    type 1 and type 2 must match the oracle and the unstaged plan."""
    import finufft_b200 as F
    rng = np.random.default_rng(91)
    rt, ct = _dt(prec)
    M, ntr = 51_111, 2
    pts = make_points(rng, dim, M, rt, "wide")[:dim]
    dpts = [cuda.from_numpy(p).cuda() for p in pts]
    for type_ in (1, 2):
        data = _rand_c(rng, (ntr, M) if type_ == 1 else (ntr,) + modes, ct)
        res = {}
        for stage in ("0", "1"):
            monkeypatch.setenv("B200_NUFFT_STAGE", stage)
            monkeypatch.setenv("B200_NUFFT_STAGE_SHIFT", "11")
            gp, op = _plans(F, oracle, type_, modes, ntr, tol, prec)
            gp.setpts(*dpts)
            res[stage] = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
            gp.destroy()
        op.setpts(*(pts[::-1] + [None] * (3 - dim)))
        want = op.execute(data)
        assert oracle.relerr(res["1"], want) <= 2 * tol
        assert oracle.relerr(res["1"], res["0"]) <= (1e-5 if prec == "f" else 1e-13)
        if type_ == 2:   # interp is order-independent: staged and unstaged agree bit for bit
            assert np.array_equal(res["1"], res["0"])


@pytest.mark.parametrize("groups", ["1", "3"])
@pytest.mark.parametrize("prec,tol", [("f", 1e-5), ("d", 1e-10)])
def test_host_api_pipelined_groups(cuda, oracle, prec, tol, groups, monkeypatch):
    """finufft[f]_execute with host arrays: the pipelined path (copy streams + events, point
    groups of consecutive user indices when B200_NUFFT_HOST_GROUPS > 1) gives the oracle's
    numbers for types 1 and 2, one and several vectors, and the adjoint."""
    import finufft_b200 as F
    monkeypatch.setenv("B200_NUFFT_HOST_GROUPS", groups)
    rng = np.random.default_rng(93)
    rt, ct = _dt(prec)
    for dim, modes in ((1, (400,)), (2, (48, 40)), (3, (20, 24, 18))):
        M = 20_011
        pts = make_points(rng, dim, M, rt, "wide")[:dim]
        for type_, ntr in ((1, 1), (2, 1), (1, 3), (2, 3)):
            hp = F.HostPlan(type_, modes, ntr, tol, 1, ct, upsampfac=2.0, allow_eps_too_small=1)
            hp.setpts(*pts)
            op = oracle.Plan(type_, list(modes[::-1]), 1, ntr, tol, rt, sigma=2.0,
                             nthr=oracle.max_threads())
            op.setpts(*(pts[::-1] + [None] * (3 - dim)))
            shape = (M,) if type_ == 1 else modes
            data = _rand_c(rng, ((ntr,) if ntr > 1 else ()) + shape, ct)
            for rep in range(2):   # a second execute reuses streams and events
                got = hp.execute(data)
                assert oracle.relerr(got, op.execute(data)) <= 2 * tol, (dim, type_, ntr, rep)
            if type_ == 1 and ntr == 1:
                fk2 = _rand_c(rng, modes, ct)
                cadj = hp.execute_adjoint(fk2)
                lhs, rhs = np.vdot(fk2, got), np.vdot(cadj, data)
                assert abs(lhs - rhs) <= (1e-4 if prec == "f" else 1e-9) * abs(lhs)
            hp.destroy()


def test_tiny_grids_wrap_inside_the_window(cuda, oracle):
    """Fine grids narrower than the sweep kernels' lane window (nf = 2*ns < 8) and shorter than
    their register rows: window columns / rows alias the same cells and must still add up."""
    import finufft_b200 as F
    rng = np.random.default_rng(97)
    for prec, tols in (("f", (3e-1, 3e-3, 1e-5)), ("d", (3e-1, 3e-3, 1e-9))):
        rt, ct = _dt(prec)
        for modes in ((1, 1), (2, 3), (3, 1), (5, 2), (2, 2, 2), (1, 4, 3)):
            dim, M = len(modes), 777
            pts = make_points(rng, dim, M, rt, "wide")[:dim]
            for tol in tols:
                for type_ in (1, 2):
                    gp, op = _plans(F, oracle, type_, modes, 1, tol, prec)
                    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
                    op.setpts(*(pts[::-1] + [None] * (3 - dim)))
                    data = _rand_c(rng, (M,) if type_ == 1 else modes, ct)
                    got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
                    want = op.execute(data)
                    assert oracle.relerr(got, want) <= max(2 * tol, 5e-6 if prec == "f" else 1e-13), \
                        (prec, modes, tol, type_)
                    gp.destroy()


def test_many_points_in_one_bin(cuda, oracle):
    """Clustered input: every point in a few bins, so bins split into many subproblems."""
    import finufft_b200 as F
    rng = np.random.default_rng(61)
    modes, M, tol = (32, 32, 32), 300_000, 1e-6
    for type_ in (1, 2):
        gp, op = _plans(F, oracle, type_, modes, 1, tol, "f", gpu_maxsubprobsize=512)
        pts = make_points(rng, 3, M, np.float32, "cluster", nf=gp.info()["nf"][::-1])
        gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts])
        op.setpts(*pts[::-1])
        assert gp.info()["nsub"] > 500
        data = _rand_c(rng, (M,) if type_ == 1 else modes, np.complex64)
        got = gp.execute(cuda.from_numpy(data).cuda()).cpu().numpy()
        want = op.execute(data)
        # 3e5 float additions per fine cell: compare with the double-precision oracle too
        assert oracle.relerr(got, want) <= 2e-5


def test_host_pointer_api(cuda, oracle):
    """finufft[f]_* entry points (host arrays), incl. execute_adjoint and error 26."""
    import finufft_b200 as F
    rng = np.random.default_rng(71)
    modes, M, tol = (30, 44), 25_000, 1e-9
    pts = make_points(rng, 2, M, np.float64)[:2]
    op1 = oracle.Plan(1, list(modes[::-1]), 1, 2, tol, np.float64, nthr=oracle.max_threads())
    op1.setpts(pts[1], pts[0])
    hp = F.HostPlan(1, modes, 2, tol, 1, "complex128", upsampfac=2.0)
    hp.setpts(*pts)
    c = _rand_c(rng, (2, M), np.complex128)
    fk = hp.execute(c)
    assert oracle.relerr(fk, op1.execute(c)) <= 2 * tol
    # adjoint of type 1 = type 2 with the opposite sign: <fk2, A c> = <A^H fk2, c>
    fk2 = _rand_c(rng, (2,) + modes, np.complex128)
    cadj = hp.execute_adjoint(fk2)
    lhs = np.vdot(fk2[0], fk[0])
    rhs = np.vdot(cadj[0], c[0])
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs)   # reference test/adjointness.cpp allows 1e-10..1e-4
    # single precision: tol below the rounding floor of the grid is an error unless allowed
    with pytest.raises(F.NufftError) as e:
        p = F.HostPlan(1, (256, 256, 256), 1, 1e-6, 1, "complex64", upsampfac=2.0)
        p.setpts(*make_points(rng, 3, 100, np.float32))
    assert e.value.code == 26
    p = F.HostPlan(1, (64, 64, 64), 1, 1e-6, 1, "complex64", upsampfac=2.0, allow_eps_too_small=1)
    p.setpts(*make_points(rng, 3, 100, np.float32))
    with pytest.raises(F.NufftError) as e:
        F.HostPlan(1, (64,), 1, 1e-6, 1, "complex128", upsampfac=0.9)
    assert e.value.code == 7


@pytest.mark.parametrize("prec,tol", [("f", 1e-5), ("d", 1e-9)])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_type3_matches_oracle_and_direct_sum(cuda, oracle, prec, tol, dim):
    """Type 3 (include/finufft/setpts.hpp:163-319, execute.hpp:432-558): device and host entry
    points against the oracle (2*tol) and the O(NM) direct sum (tolsweep-style threshold)."""
    import finufft_b200 as F
    rng = np.random.default_rng(40 + dim)
    rt, ct = _dt(prec)
    M, N, ntr = 3000, 2500, 2
    pts = [rng.uniform(-np.pi, np.pi, M).astype(rt) + rt(0.3 * d) for d in range(dim)]
    frq = [(rng.uniform(-20, 20, N) + 3.0 * (d + 1)).astype(rt) for d in range(dim)]
    c = _rand_c(rng, (ntr, M), ct)
    gp = F.Plan(3, dim, ntr, tol, 1, ct, upsampfac=2.0)
    gp.setpts(*[cuda.from_numpy(a).cuda() for a in pts],
              **dict(zip("stu", [cuda.from_numpy(a).cuda() for a in frq])))
    got = gp.execute(cuda.from_numpy(c).cuda()).cpu().numpy()
    op = oracle.Plan(3, [1] * dim, 1, ntr, tol, rt, sigma=2.0, dim=dim, nthr=oracle.max_threads())
    lp, lf = pts[::-1] + [None] * (3 - dim), frq[::-1] + [None] * (3 - dim)
    op.setpts(lp[0], lp[1], lp[2], lf[0], lf[1], lf[2])
    want = op.execute(c)
    assert oracle.relerr(got.reshape(-1), want.reshape(-1)) <= 2 * tol
    for b in range(ntr):
        ds = oracle.dirft(3, lp[0], lp[1], lp[2], c[b].astype(np.complex128), 1,
                          s=lf[0], t=lf[1], u=lf[2])
        assert oracle.relerr(got[b], ds) <= 10 * tol
    hp = F.HostPlan(3, dim, ntr, tol, 1, ct, upsampfac=2.0, allow_eps_too_small=1)
    hp.setpts(*pts, **dict(zip("stu", frq)))
    hgot = hp.execute(c)
    assert oracle.relerr(hgot.reshape(-1), got.reshape(-1)) <= (1e-5 if prec == "f" else 1e-12)
    # adjoint of type 3 (execute.hpp:515-545): targets -> sources with the conjugate kernel;
    # checked against the direct sum sum_k F_k exp(-i s_k.x_j) and the identity <F, A c> = <A^H F, c>
    F_in = _rand_c(rng, (ntr, N), ct)
    adj = hp.execute_adjoint(F_in)
    assert adj.shape == (ntr, M)
    for b in range(ntr):
        ds = oracle.dirft(3, lf[0], lf[1], lf[2], F_in[b].astype(np.complex128), -1,
                          s=lp[0], t=lp[1], u=lp[2])
        assert oracle.relerr(adj[b], ds) <= 10 * tol
        lhs = np.vdot(F_in[b].astype(np.complex128), hgot[b].astype(np.complex128))
        rhs = np.vdot(adj[b].astype(np.complex128), c[b].astype(np.complex128))
        assert abs(lhs - rhs) <= (1e-4 if prec == "f" else 1e-10) * abs(lhs)
    gp.destroy()
    hp.destroy()


def test_error_contract_gpu(cuda):
    """reference test/cuda/cufinufft_error_handling.cu / test_makeplan.c / multigpu test."""
    import finufft_b200 as F
    lib = F.load()
    p = C.c_void_p()
    nm = (C.c_int64 * 3)(16, 16, 16)
    o = F._lib.CufinufftOpts()
    lib.cufinufft_default_opts(C.byref(o))
    o.gpu_device_id = 1234
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 1, C.c_float(1e-4), C.byref(p), C.byref(o)) == 15
    o.gpu_device_id = 0
    # nonstandard sigma: error 8 with the default gpu_kerevalmeth = 1, error 7 (sigma <= 1) with
    # gpu_kerevalmeth = 0 (reference src/cuda/makeplan.cu:60-72, test/cuda/test_makeplan.c:175-185)
    o.upsampfac = 1.0
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 1, C.c_float(1e-4), C.byref(p), C.byref(o)) == 8
    o.gpu_kerevalmeth = 0
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 1, C.c_float(1e-4), C.byref(p), C.byref(o)) == 7
    o.gpu_kerevalmeth = 1
    o.upsampfac = 2.0
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 1, C.c_float(1e-4), C.byref(p), C.byref(o)) == 0
    before = cuda.cuda.current_device()
    assert lib.cufinufftf_setpts(p, 2 ** 31, None, None, None, 0, None, None, None) == 14
    assert lib.cufinufft_destroy(p) == 16     # wrong precision handle
    assert lib.cufinufftf_destroy(p) == 0
    assert cuda.cuda.current_device() == before
    # fine grid below 2*ns in one dimension -> error 3 at setpts (CPU spreadcheck semantics)
    sp = F.Plan(1, (6, 40), 1, 1e-6, 1, "complex64", gpu_spreadinterponly=1)
    with pytest.raises(F.NufftError) as e:
        sp.setpts(cuda.zeros(4, device="cuda"), cuda.zeros(4, device="cuda"))
    assert e.value.code == 3


def test_user_stream_and_repeat_setpts(cuda, oracle):
    """Work is issued on opts.gpu_stream; setpts may be called again with a different M."""
    import finufft_b200 as F
    rng = np.random.default_rng(81)
    st = cuda.cuda.Stream()
    modes, tol = (36, 28, 20), 1e-5
    gp = F.Plan(1, modes, 1, tol, 1, "complex64", upsampfac=2.0, gpu_stream=st.cuda_stream)
    op = oracle.Plan(1, list(modes[::-1]), 1, 1, tol, np.float32, nthr=oracle.max_threads())
    for M in (50_000, 7_000, 90_000):
        pts = make_points(rng, 3, M, np.float32)
        with cuda.cuda.stream(st):
            d = [cuda.from_numpy(p).cuda() for p in pts]
            c = _rand_c(rng, (M,), np.complex64)
            dc = cuda.from_numpy(c).cuda()
        st.synchronize()
        gp.setpts(*d)
        out = gp.execute(dc)
        st.synchronize()
        op.setpts(*pts[::-1])
        assert oracle.relerr(out.cpu().numpy(), op.execute(c)) <= 2 * tol


# ----------------------------------------------------------------------------- full size
@pytest.mark.parametrize("workload", ["c3", "c2"])
def test_full_size_properties(cuda, oracle, workload):
    """BASELINE sizes (3D 256^3 / 2D 2048^2, M = 1e8 single precision): properties that need no
    oracle run: the permutation is a stable sort of the oracle's bin keys (bins recomputed on a
    strided sample by the oracle), type-1/type-2 adjointness, linearity, and spot modes against
    the direct sum."""
    import finufft_b200 as F
    free, _ = cuda.cuda.mem_get_info()
    if free < 30e9:
        pytest.skip("needs ~30 GB of device memory")
    modes, tol = ((256, 256, 256), 1e-6) if workload == "c3" else ((2048, 2048), 1e-5)
    dim, M = len(modes), 100_000_000
    g = cuda.Generator(device="cuda").manual_seed(5)
    pts = [(cuda.rand(M, device="cuda", generator=g) * 2 - 1) * np.pi for _ in range(dim)]
    p1 = F.Plan(1, modes, 1, tol, 1, "complex64", upsampfac=2.0)
    p1.setpts(*pts)
    nf = p1.info()["nf"]
    perm = cuda.from_numpy(p1.sort_permutation().astype(np.int64)).cuda()
    # (a) it is a permutation
    chk = cuda.zeros(M, dtype=cuda.int32, device="cuda")
    chk[perm] = 1
    assert int(chk.sum()) == M
    del chk
    # (b) keys along the permutation are non-decreasing and ties ascend in index: verify on
    #     windows of consecutive sorted positions with the oracle's bin function
    host_pts = [p.cpu().numpy() for p in pts]
    lp = host_pts[::-1] + [None] * (3 - dim)
    hperm = perm.cpu().numpy()
    for start in (0, 12_345_678, 50_000_000, M - 200_000):
        idx = hperm[start:start + 200_000]
        sub = [None if a is None else np.ascontiguousarray(a[idx]) for a in lp]
        _, bins = oracle.bin_sort(sub[0], sub[1], sub[2], nf)
        assert np.all(np.diff(bins) >= 0)
        same = np.diff(bins) == 0
        assert np.all(np.diff(idx)[same] > 0)
    # (c) adjointness <F, A c> = <A^H F, c> with A = type 1 (+i), A^H = type 2 (-i)
    c = cuda.view_as_complex(cuda.randn(M, 2, device="cuda", generator=g))
    fk = p1.execute(c)
    p2 = F.Plan(2, modes, 1, tol, -1, "complex64", upsampfac=2.0)
    p2.setpts(*pts)
    Fm = cuda.view_as_complex(cuda.randn(*modes, 2, device="cuda", generator=g))
    c2 = p2.execute(Fm)
    lhs = cuda.vdot(Fm.reshape(-1).to(cuda.complex128), fk.reshape(-1).to(cuda.complex128))
    rhs = cuda.vdot(c2.to(cuda.complex128), c.to(cuda.complex128))
    assert abs(complex(lhs) - complex(rhs)) <= 1e-4 * abs(complex(lhs))
    # (d) linearity: A(2c) = 2 A(c) exactly in floating point up to accumulation order
    fk2 = p1.execute(2 * c)
    assert float((fk2 - 2 * fk).abs().max()) <= 1e-3 * float(fk.abs().max())
    # (e) a few modes against the double-precision direct sum (reference test/finufft3d_test.cpp:82-95)
    hc = c.cpu().numpy().astype(np.complex128)
    hfk = fk.cpu().numpy()
    rng = np.random.default_rng(1)
    worst = 0.0
    scale = float(np.sqrt(np.mean(np.abs(hfk) ** 2)))
    for _ in range(3):
        k = [int(rng.integers(-(m // 2), (m - 1) // 2 + 1)) for m in modes]
        ph = sum(float(kk) * a.astype(np.float64) for kk, a in zip(k, host_pts))
        exact = np.sum(hc * np.exp(1j * ph))
        got = hfk[tuple(kk + m // 2 for kk, m in zip(k, modes))]
        worst = max(worst, abs(got - exact) / scale)
    # single precision cannot beat the rounding floor eps_round = 0.48*eps*max(nf) of the fine
    # grid (reference include/finufft/setpts.hpp:29-53; it is why the CPU library needs
    # allow_eps_too_small for this config): 2.9e-5 at nf=512, 2.3e-4 at nf=4096
    assert worst <= max(10 * tol, 0.48 * np.finfo(np.float32).eps * max(nf))
