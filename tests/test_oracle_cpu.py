"""CPU-only: pin the oracle (oracle/finufft_oracle.cpp) against
 (1) the reference's own known-answer table test/testutils.cpp:39-54 (next235even_true),
 (2) the reference's own src/common sources compiled into oracle/_ref (when present),
 (3) committed golden vectors generated from oracle/_ref (tests/golden/make_golden.py),
 (4) the reference tests' accuracy criterion: direct sums with test/tolsweep.cpp thresholds,
     the spread/interp duality of test/spreadinterp1d_test.cpp and test/adjointness.cpp.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

# reference test/testutils.cpp:39-54: next235even(n) for n = 0..99
NEXT235EVEN_TRUE = [
    2, 2, 2, 4, 4, 6, 6, 8, 8, 10, 10, 12, 12, 16, 16, 16, 16, 18, 18, 20, 20, 24, 24, 24, 24,
    30, 30, 30, 30, 30, 30, 32, 32, 36, 36, 36, 36, 40, 40, 40, 40, 48, 48, 48, 48, 48, 48, 48,
    48, 50, 50, 54, 54, 54, 54, 60, 60, 60, 60, 60, 60, 64, 64, 64, 64, 72, 72, 72, 72, 72, 72,
    72, 72, 80, 80, 80, 80, 80, 80, 80, 80, 90, 90, 90, 90, 90, 90, 90, 90, 90, 90, 96, 96, 96,
    96, 96, 96, 100, 100, 100]


def test_next235_known_answers(oracle):
    assert [oracle.next235(n, 2) for n in range(100)] == NEXT235EVEN_TRUE


def test_gaussquad_integrates_polynomials(oracle):
    # reference test/testutils.cpp:70-87: n-point rule exact to degree 2n-1
    for n in (2, 7, 16, 40):
        x, w = oracle.gaussquad(n)
        for deg in range(0, 2 * n, 2):
            assert abs(np.sum(w * x ** deg) - 2.0 / (deg + 1)) < 1e-13


def test_golden_vectors(oracle):
    g = json.load(open(os.path.join(HERE, "golden", "plan_math.json")))
    for case in g["kernel"]:
        err, ns, beta, _ = oracle.kernel_setup(case["tol"], case["dim"], case["type"],
                                               case["sigma"],
                                               np.float32 if case["is_float"] else np.float64, True)
        assert err == 0 and ns == case["ns"]
        assert abs(beta - case["beta"]) < 1e-13
    for case in g["pswf"]:
        got = oracle.pswf(case["c"], np.array(case["x"]))
        assert np.max(np.abs(got - np.array(case["psi"]))) < 1e-15
    for case in g["polyfit"]:
        dt = np.float32 if case["dtype"] == "f32" else np.float64
        got = oracle.polyfit_pswf(case["ns"], case["beta"], case["panel"], case["n"], dt)
        want = np.array(case["coef"], dtype=dt)
        assert np.array_equal(got, want), (case["dtype"], case["ns"], case["panel"])
    for case in g["lowest_sigma"]:
        got = oracle.lowest_sigma(*case["args"])
        assert abs(got - case["value"]) < 1e-14
    for case in g["nhg_type3"]:
        nf, h, gam = oracle.nhg_type3(*case["args"])
        assert nf == case["nf"] and abs(h - case["h"]) < 1e-15 and abs(gam - case["gam"]) < 1e-12


def test_oracle_matches_reference_common(oracle):
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    assert all(oracle.next235(n, 2) == R.ref_next235(n, 2) for n in range(3000))
    x = np.linspace(-1, 1, 257)
    out = np.zeros_like(x)
    for c in (2.3, 9.99, 14.0872, 16.4434, 23.5119, 31.0, 44.0):
        R.ref_pswf(C.c_double(c), C.c_int64(x.size), x.ctypes.data_as(C.c_void_p),
                   out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, oracle.pswf(c, x))
    for tol, dim, typ, sig, isf in [(1e-6, 3, 1, 2.0, 1), (1e-5, 2, 2, 2.0, 1), (1e-9, 2, 1, 2.0, 0),
                                    (1e-9, 1, 1, 2.0, 0), (1e-6, 3, 3, 1.25, 1), (1e-3, 1, 1, 1.25, 1),
                                    (1e-12, 3, 2, 2.0, 0), (1e-14, 1, 1, 1.5, 0)]:
        ns, beta = C.c_int(), C.c_double()
        R.ref_kernel_ns_beta(C.c_double(tol), dim, typ, C.c_double(sig), isf, C.byref(ns),
                             C.byref(beta))
        err, ns2, beta2, _ = oracle.kernel_setup(tol, dim, typ, sig,
                                                 np.float32 if isf else np.float64, True)
        assert err == 0 and ns2 == ns.value and beta2 == beta.value
    for suf, dt in (("f32", np.float32), ("f64", np.float64)):
        for ns_, beta_ in ((4, 9.3748), (7, 16.4434), (12, 28.2243)):
            n = min(19, ns_ + 3)
            for panel in range(ns_):
                b = np.zeros(n, dtype=dt)
                getattr(R, "ref_polyfit_pswf_" + suf)(ns_, C.c_double(beta_), panel, n,
                                                      b.ctypes.data_as(C.c_void_p))
                assert np.array_equal(oracle.polyfit_pswf(ns_, beta_, panel, n, dt), b)
    xg, wg = oracle.gaussquad(26)
    xr, wr = np.zeros(26), np.zeros(26)
    R.ref_gaussquad(26, xr.ctypes.data_as(C.c_void_p), wr.ctypes.data_as(C.c_void_p))
    assert np.array_equal(xg, xr) and np.array_equal(wg, wr)


def test_config_table(oracle):
    """ns / nc / nf of the BASELINE configs (computed there with the reference's
    own sources)."""
    want = {  # (tol, dim, type, dtype): (ns, nc)
        (1e-6, 3, 1, np.float32): (7, 10), (1e-5, 2, 2, np.float32): (6, 9),
        (1e-9, 2, 1, np.float64): (10, 13), (1e-9, 1, 1, np.float64): (10, 13),
    }
    for (tol, dim, typ, dt), (ns, nc) in want.items():
        err, ns2, beta, tolu = oracle.kernel_setup(tol, dim, typ, 2.0, dt, True)
        coef, nc2 = oracle.horner(ns2, beta, tolu, dt)
        assert (ns2, nc2) == (ns, nc)
    p = oracle.Plan(1, [256, 256, 256], 1, 1, 1e-6, np.float32)
    assert p.nf == [512, 512, 512] and abs(p.beta - 16.4434) < 1e-4
    p = oracle.Plan(2, [2048, 2048], 1, 1, 1e-5, np.float32)
    assert p.nf == [4096, 4096]
    p = oracle.Plan(1, [1000000], 1, 1, 1e-9, np.float64)
    assert p.nf == [2000000]


# test/tolsweep.cpp:36,53-56: pass iff relerr <= max(floor, slack*tol)
FLOOR = {np.float32: (2e-5, 2e-5, 1e-5), np.float64: (3e-14, 3e-14, 3e-14)}
SLACK = (4, 4, 5)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("dim,ms", [(1, [50]), (2, [25, 40]), (3, [10, 11, 12])])
@pytest.mark.parametrize("type_", [1, 2])
def test_oracle_vs_direct_sum_tolsweep(oracle, dt, dim, ms, type_):
    rng = np.random.default_rng(3)
    M = 500
    ct = np.complex64 if dt == np.float32 else np.complex128
    tols = (1e-2, 1e-4, 1e-6) if dt == np.float32 else (1e-3, 1e-6, 1e-9, 1e-12)
    pts = [rng.uniform(-np.pi, np.pi, M).astype(dt) for _ in range(dim)] + [None] * (3 - dim)
    for tol in tols:
        p = oracle.Plan(type_, ms, 1, 1, tol, dt, sigma=2.0, nthr=2)
        p.setpts(*pts)
        n_in = M if type_ == 1 else int(np.prod(ms))
        data = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(ct)
        got = p.execute(data)
        want = oracle.dirft(type_, pts[0], pts[1], pts[2], data, 1, n_modes=ms)
        assert oracle.relerr(got, want) <= max(FLOOR[dt][dim - 1], SLACK[dim - 1] * tol)


def test_oracle_type3_vs_direct_sum(oracle):
    rng = np.random.default_rng(4)
    for dim in (1, 2, 3):
        M, N = 300, 250
        pts = [rng.uniform(-np.pi, np.pi, M) for _ in range(dim)] + [None] * (3 - dim)
        frq = [rng.uniform(-15, 15, N) + 2.0 for _ in range(dim)] + [None] * (3 - dim)
        p = oracle.Plan(3, [1] * dim, 1, 1, 1e-9, np.float64, dim=dim)
        p.setpts(pts[0], pts[1], pts[2], frq[0], frq[1], frq[2])
        c = rng.standard_normal(M) + 1j * rng.standard_normal(M)
        got = p.execute(c)
        want = oracle.dirft(3, pts[0], pts[1], pts[2], c, 1, s=frq[0], t=frq[1], u=frq[2])
        assert oracle.relerr(got, want) < 5e-9


def test_oracle_spread_interp_duality(oracle):
    """test/adjointness.cpp:27-45 / test/spreadinterp1d_test.cpp:104-160: <g, S c> = <I g, c>,
    and interpolating the all-ones grid returns the kernel sum at every point."""
    rng = np.random.default_rng(5)
    err, ns, beta, tolu = oracle.kernel_setup(1e-6, 1, 1, 2.0, np.float64, True)
    coef, nc = oracle.horner(ns, beta, tolu, np.float64)
    nf, M = [120], 400
    x = rng.uniform(-np.pi, np.pi, M)
    c = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    perm, _ = oracle.bin_sort(x, None, None, nf)
    fw = oracle.spread(nf, x, None, None, c, perm, coef)
    g = rng.standard_normal(nf[0]) + 1j * rng.standard_normal(nf[0])
    ci = oracle.interp(nf, x, None, None, g, perm, coef)
    lhs, rhs = np.vdot(g, fw), np.vdot(ci, c)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    ones = oracle.interp(nf, x, None, None, np.ones(nf[0], dtype=np.complex128), perm, coef)
    # kernel sum is constant in the offset up to the aliasing error ~tol
    assert np.ptp(ones.real) < 1e-5 * np.mean(ones.real) and np.max(np.abs(ones.imag)) == 0
    # total mass: sum(fw) = sum_j c_j * kersum_j
    assert abs(np.sum(fw) - np.sum(c * ones.real)) < 1e-10 * np.abs(np.sum(c * ones.real)) + 1e-9


def test_oracle_sort_is_stable_counting_sort(oracle):
    """include/finufft/spread.hpp:459-584: ascending bin, ties in ascending index; bins from
    trunc(fold_rescale(x,N) / binsize) with 16 x 4 x 4 cells per bin."""
    rng = np.random.default_rng(6)
    for dt in (np.float32, np.float64):
        x, y, z = [rng.uniform(-10, 10, 5000).astype(dt) for _ in range(3)]
        nf = [60, 36, 48]
        perm, bins = oracle.bin_sort(x, y, z, nf)
        assert np.array_equal(perm, np.argsort(bins, kind="stable"))
        nb1, nb2 = int(dt(nf[0]) / 16 + 1), int(dt(nf[1]) / 4 + 1)
        i1 = (oracle.fold_rescale(x, nf[0]) * dt(1 / 16)).astype(np.int64)
        i2 = (oracle.fold_rescale(y, nf[1]) * dt(1 / 4)).astype(np.int64)
        i3 = (oracle.fold_rescale(z, nf[2]) * dt(1 / 4)).astype(np.int64)
        assert np.array_equal(bins, i1 + nb1 * (i2 + nb2 * i3))


def test_oracle_error_codes(oracle):
    for args, code in [((1e-6, 3, 1, 1.0, np.float32, False), 7),
                       ((1e-9, 3, 1, 2.0, np.float32, False), 26),
                       ((1e-20, 1, 1, 2.0, np.float64, False), 26)]:
        assert oracle.kernel_setup(*args)[0] == code
    assert oracle.kernel_setup(1e-9, 3, 1, 2.0, np.float32, True)[0] == 0


def test_direct_sum_checker_is_the_reference_code(oracle):
    """The direct-sum checker used by every accuracy test is the reference's own
    test/utils/dirft{1,2,3}d.hpp + norms.hpp compiled into oracle/_ref; our restatement of them
    (orc_dirft*) must agree with it to rounding, for every type and dimension."""
    if oracle.ref_dirft_lib() is None:
        pytest.skip("oracle/_ref/libfinufft_ref_dirft.so not present (no /root/reference)")
    rng = np.random.default_rng(3)
    M, nk = 300, 170
    modes = {1: [37], 2: [12, 9], 3: [6, 7, 5]}
    for dim in (1, 2, 3):
        pts = [rng.uniform(-np.pi, np.pi, M) for _ in range(dim)] + [None] * (3 - dim)
        frq = [rng.uniform(-30, 30, nk) for _ in range(dim)] + [None] * (3 - dim)
        N = int(np.prod(modes[dim]))
        c = rng.standard_normal(M) + 1j * rng.standard_normal(M)
        f = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for sign in (+1, -1):
            a = oracle.dirft(1, *pts, c, sign, n_modes=modes[dim], impl="ref")
            b = oracle.dirft(1, *pts, c, sign, n_modes=modes[dim], impl="port")
            assert oracle.relerr(b, a) < 1e-13
            # one thread = the reference routine called once, untouched by the OpenMP driver
            a1 = oracle.dirft(1, *pts, c, sign, n_modes=modes[dim], impl="ref", nthr=1)
            assert oracle.relerr(a, a1) < 1e-13
            a = oracle.dirft(2, *pts, f, sign, n_modes=modes[dim], impl="ref")
            b = oracle.dirft(2, *pts, f, sign, n_modes=modes[dim], impl="port")
            assert oracle.relerr(b, a) < 1e-13
            a = oracle.dirft(3, *pts, c, sign, s=frq[0], t=frq[1], u=frq[2], impl="ref")
            b = oracle.dirft(3, *pts, c, sign, s=frq[0], t=frq[1], u=frq[2], impl="port")
            assert oracle.relerr(b, a) < 1e-12
    # relerr itself is the reference's relerrtwonorm
    u = rng.standard_normal(50) + 1j * rng.standard_normal(50)
    v = u + 1e-3 * rng.standard_normal(50)
    assert abs(oracle.relerr(v, u) - np.linalg.norm(v - u) / np.linalg.norm(u)) < 1e-15
