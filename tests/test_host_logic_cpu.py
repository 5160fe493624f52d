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
