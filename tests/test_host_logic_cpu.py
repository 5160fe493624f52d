"""CPU-only: the product's own host-side plan mathematics (finufft_b200/csrc/planmath.cpp,
reached through the b200_host_* entry points) against the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    from finufft_b200 import hostmath
    return hostmath


CASES = [(1e-6, 3, 1, 2.0, np.float32), (1e-5, 2, 2, 2.0, np.float32), (1e-9, 2, 1, 2.0, np.float64),
         (1e-9, 1, 1, 2.0, np.float64), (1e-6, 3, 3, 1.25, np.float32), (1e-3, 2, 1, 2.0, np.float32),
         (1e-12, 3, 1, 2.0, np.float64), (1e-2, 1, 2, 1.25, np.float64), (1e-14, 1, 1, 2.0, np.float64)]


@pytest.mark.parametrize("tol,dim,typ,sigma,dt", CASES)
def test_kernel_choice_and_table(hm, oracle, tol, dim, typ, sigma, dt):
    err, ns, beta, tab = hm.kernel(tol, dim, typ, sigma, dt)
    oerr, ons, obeta, otol = oracle.kernel_setup(tol, dim, typ, sigma, dt, True)
    assert err == oerr == 0 and ns == ons and abs(beta - obeta) < 1e-13
    ocoef, onc = oracle.horner(ons, obeta, otol, dt)
    assert tab.shape == (onc, ons)
    # same kernel FUNCTION: compare evaluated window values over the support (coefficients of
    # high-degree fits are ill-conditioned, the values are not)
    xs = np.linspace(-ns / 2 + 1e-3, -ns / 2 + 1 - 1e-3, 23)
    ours = np.array([oracle.eval_stencil(dt(x), tab) for x in xs], dtype=np.float64)
    ref = np.array([oracle.eval_stencil(dt(x), ocoef) for x in xs], dtype=np.float64)
    eps = np.finfo(dt).eps
    assert np.max(np.abs(ours - ref)) <= 64 * eps
    if dt == np.float32 and ns <= 10:
        assert np.array_equal(tab, ocoef)


def test_error_codes(hm):
    assert hm.kernel(1e-6, 3, 1, 1.0, np.float32)[0] == 7
    assert hm.kernel(1e-9, 3, 1, 2.0, np.float32, allow_small=False)[0] == 26
    assert hm.kernel(1e-9, 3, 1, 2.0, np.float32, allow_small=True)[0] == 0
    assert hm.kernel(1e-20, 1, 1, 2.0, np.float64, allow_small=False)[0] == 26


def test_fine_grid_sizes(hm, oracle):
    for sigma in (2.0, 1.25):
        for ns in (2, 7, 16):
            for modes in list(range(1, 200)) + [256, 2048, 1000000, 12345]:
                want = oracle.next235(max(int(np.ceil(sigma * modes)), 2 * ns), 2)
                assert hm.fine_grid(sigma, modes, ns) == want
    assert hm.fine_grid(2.0, 256, 7) == 512 and hm.fine_grid(2.0, 2048, 6) == 4096
    assert hm.fine_grid(2.0, 10 ** 12, 7) == -1


@pytest.mark.parametrize("dt,tol", [(np.float32, 1e-6), (np.float32, 1e-3), (np.float64, 1e-9)])
def test_fseries_matches_oracle(hm, oracle, dt, tol):
    """The deconvolution factors follow the reference's working-precision phase winding
    (include/finufft/makeplan.hpp:72-105), so they agree with the oracle to the last bits."""
    err, ns, beta, tab = hm.kernel(tol, 1, 1, 2.0, dt)
    for nf in (2 * ns, 90, 512, 4096):
        ours = hm.fseries(nf, tab)
        ref = oracle.fseries(nf, tab)
        assert ours.dtype == ref.dtype
        assert np.max(np.abs(ours.astype(np.float64) - ref.astype(np.float64))) <= \
            4 * np.finfo(dt).eps * abs(float(ref[0]))
        assert ours[0] > 0 and np.all(np.sign(ours[: nf // 4]) == (-1.0) ** np.arange(nf // 4))


def test_auto_upsampfac_search_matches_reference():
    """Automatic upsampfac (host API, upsampfac = 0): the smallest feasible sigma and the
    feasibility test against the reference's own src/common/kernel.cpp:203-257 compiled into
    oracle/_ref (analytic_upsampfac, upsampfac_feasible), the two probe values of SURVEY.md 8(c),
    and sanity of the B200 cost-model pick (dense point sets keep sigma = 2, sparse ones on a large
    grid go below it and stay feasible)."""
    import ctypes as C
    import finufft_b200
    lib = finufft_b200.load()
    assert abs(lib.b200_host_smallest_sigma(1e-9, 1, 1, 0, 1e6) - 1.8394) < 2e-4
    assert abs(lib.b200_host_smallest_sigma(1e-9, 2, 1, 0, 512.0) - 1.2027) < 2e-4
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libfinufft_ref_common.so")
    if os.path.exists(ref_path):
        ref = C.CDLL(ref_path)
        ref.ref_analytic_upsampfac.restype = C.c_double
        ref.ref_analytic_upsampfac.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double]
        ref.ref_upsampfac_feasible.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                               C.c_double]
        for is_float, tols in ((1, (1e-2, 1e-3, 1e-4, 1e-5, 1e-6)), (0, (1e-3, 1e-6, 1e-9, 1e-12, 1e-14))):
            for tol in tols:
                for dim in (1, 2, 3):
                    for type_ in (1, 2):
                        for N in (16.0, 200.0, 4096.0, 1e6):
                            a = lib.b200_host_smallest_sigma(tol, dim, type_, is_float, N)
                            b = ref.ref_analytic_upsampfac(tol, dim, type_, is_float, N)
                            assert abs(a - b) < 1e-9, (is_float, tol, dim, type_, N, a, b)
                            for s in (1.15, 1.25, 1.5, 2.0, 2.5):
                                assert lib.b200_host_sigma_feasible(s, tol, dim, type_, is_float, N) == \
                                    ref.ref_upsampfac_feasible(s, tol, dim, type_, is_float, N)
    modes = (C.c_int64 * 3)(256, 256, 256)
    assert lib.b200_host_choose_sigma(1e-6, 3, 1, 1, modes, 1e8) == 2.0      # C3: spread-dominated
    s = lib.b200_host_choose_sigma(1e-4, 3, 1, 1, modes, 1e4)                # 1e4 points on 512^3
    assert 1.15 <= s < 2.0 and lib.b200_host_sigma_feasible(s, 1e-4, 3, 1, 1, 256.0)
    modes2 = (C.c_int64 * 3)(512, 512, 1)
    s = lib.b200_host_choose_sigma(1e-9, 2, 1, 0, modes2, 2e3)
    assert 1.2 <= s < 2.0 and lib.b200_host_sigma_feasible(s, 1e-9, 2, 1, 0, 512.0)
    assert lib.b200_host_choose_sigma(1e-9, 2, 1, 0, modes2, 1e7) == 2.0     # C4


def test_sigma_candidates_are_the_reference_minimisers(tmp_path):
    """The (sigma, ns) candidates the automatic-upsampfac search scores, for types 1, 2 and 3,
    against a trace of the reference's own minimiser (include/finufft/heuristics.hpp:82-107,
    called through oracle/_ref with a cost functor that records its arguments): same count, same
    order, sigma equal to 1e-9.  The widths agree except where the reference evaluates its width
    law exactly on the edge of its ceil() (its candidates are built to sit there) and rounding
    puts it one wider; ours sit one part in 1e12 above the edge and get the intended width."""
    import ctypes as C
    import itertools
    import finufft_b200
    lib = finufft_b200.load()
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libfinufft_ref.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref/libfinufft_ref.so not built (no /root/reference)")
    ref = C.CDLL(ref_path)
    ref.ref_minimize_trace.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                       C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    ncases = nlists = 0
    for tol, dim, typ, isf, N in itertools.product(
            (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-9, 1e-12, 1e-14), (1, 2, 3), (1, 2, 3), (0, 1),
            (1.0, 64.0, 512.0, 1e4, 1e6)):
        if (typ == 3) != (N == 1.0) or (isf and tol < 1e-6):
            continue
        if isf:
            tol = float(np.float32(tol))    # the plan holds tol in its own precision
        a, an = (C.c_double * 32)(), (C.c_int * 32)()
        b, bn = (C.c_double * 32)(), (C.c_int * 32)()
        na = lib.b200_host_sigma_candidates(tol, dim, typ, isf, N, 2.5, a, an, 32)
        nb = ref.ref_minimize_trace(tol, dim, typ, isf, N, b, bn, 32)
        assert na == nb, (tol, dim, typ, isf, N, list(a[:na]), list(b[:nb]))
        for i in range(na):
            assert abs(a[i] - b[i]) < 1e-9 and bn[i] - an[i] in (0, 1), \
                (tol, dim, typ, isf, N, i, a[i], an[i], b[i], bn[i])
        assert an[0] == bn[0]
        ncases += 1
        nlists += na
    assert ncases > 300 and nlists > 2 * ncases


def test_type3_auto_upsampfac_choice():
    """Type 3 with finufft_opts.upsampfac = 0: sigma3 is picked at setpts from the half-widths
    and the point counts (include/finufft/setpts.hpp:186-200, heuristics.hpp:130-150) with this
    device's cost model.  Checks of the pick itself: always a member of the candidate set or 2.0,
    feasible; few points on a wide grid (FFT-dominated) go below 2; many points on a small grid
    (spread-dominated) keep 2."""
    import ctypes as C
    import finufft_b200
    lib = finufft_b200.load()
    D3 = C.c_double * 3
    for isf, tol in ((1, 1e-5), (0, 1e-9), (0, 1e-12)):
        for dim in (1, 2, 3):
            for M, N, X, S in ((1e3, 1e3, 3.0, 400.0), (1e7, 1e7, 3.14, 100.0), (1e8, 1e3, 1.0, 8.0)):
                s = lib.b200_host_choose_sigma_type3(tol, dim, isf, M, N, D3(X, X, X), D3(S, S, S))
                cs, cn = (C.c_double * 32)(), (C.c_int * 32)()
                t = float(np.float32(tol)) if isf else tol
                n = lib.b200_host_sigma_candidates(t, dim, 3, isf, 1.0, 2.0, cs, cn, 32)
                assert s == 2.0 or any(abs(s - cs[i]) < 1e-12 for i in range(n)), (isf, tol, dim, s)
                assert lib.b200_host_sigma_feasible(s, tol, dim, 3, isf, 1.0)
    # 1e3 points against a 3D grid of ~ (2*2*400*3/pi)^3 cells: the FFT is everything
    assert lib.b200_host_choose_sigma_type3(1e-9, 3, 0, 1e3, 1e3, D3(3, 3, 3), D3(400, 400, 400)) < 1.5
    # 1e8 sources spread onto a grid of a few dozen cells per dimension: the spread is everything
    assert lib.b200_host_choose_sigma_type3(1e-5, 3, 1, 1e8, 1e3, D3(1, 1, 1), D3(8, 8, 8)) == 2.0


def test_bench_input_streams_are_the_reference_perftest_streams(tmp_path):
    """tools/perfdata.py (bench inputs) against the reference's own generator compiled from where
    it lies (perftest/randunif.h:29-73), when /root/reference is present; always: determinism,
    independence of the chunk offset, value range."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import perfdata
    a = np.empty(200_000, dtype=np.float32)
    perfdata.fill(a, "X", np.float32(np.pi), 0.0)
    b = np.empty(70_000, dtype=np.float32)
    perfdata.fill(b, "X", np.float32(np.pi), 0.0, first=65_000)
    assert np.array_equal(a[65_000:135_000], b)
    assert a.min() >= -np.pi and a.max() <= np.pi and abs(a.mean()) < 0.02
    d = np.empty(1000, dtype=np.float64)
    perfdata.fill(d, "C")
    assert np.all(np.abs(d) < 1.0)
    hdr = "/root/reference/perftest/randunif.h"
    if not os.path.exists(hdr):
        return
    src = tmp_path / "ru.cpp"
    src.write_text('#include "%s"\n#include <cstdio>\nint main(){ std::vector<float> v(200000);'
                   ' perftest_rand::fill<float>(v.data(), (std::int64_t)v.size(), perftest_rand::X,'
                   ' 3.14159265358979323846f, 0.f);'
                   ' fwrite(v.data(), 4, v.size(), stdout); }\n' % hdr)
    exe = tmp_path / "ru"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", str(src), "-o", str(exe)])
    ref = np.frombuffer(subprocess.check_output([str(exe)]), dtype=np.float32)
    assert np.array_equal(ref, a)
