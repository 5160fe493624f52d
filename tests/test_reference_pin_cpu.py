"""Pins the CPU oracle (oracle/finufft_oracle.cpp, our restatement) to the REFERENCE library's
own output.

tests/golden/reference_vectors.npz was produced by oracle/_ref/libfinufft_ref.so = the
reference's src/*.cpp + include/finufft/*.hpp compiled where they lie (oracle/build.py::
build_ref_library; third-party xsimd / POET / FFTW replaced by the stand-ins in oracle/shim/).
Here the restatement must reproduce, on the same seeded inputs:
  * the reference's sort permutation, bit for bit (spread.hpp:459-584),
  * its kernel width, polynomial degree and fine grid (makeplan.hpp, kernel.cpp),
  * its outputs to rounding: 1e-13 relative l2 in double, 5e-6 in single (two single-precision
    pipelines with different Horner / accumulation / FFT orders; measured 3e-7..1.3e-6).
When the compiled reference is present (this container, and the GPU box via the snapshot) the
golden file itself is re-derived from it and larger live comparisons run as well.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_vectors as G  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))


def _plan(cls, name, nthr=1):
    type_, modes, M, tol, prec, ntr, modeord, sigma, kind, N = G.CASES[name]
    rt = np.float32 if prec == "f32" else np.float64
    pts, frq, data = G.case_inputs(name)
    dim = len(pts)
    p = cls(type_, list(modes) if type_ != 3 else [1] * dim, 1, ntr, tol, rt, sigma=sigma,
            modeord=modeord, nthr=nthr, dim=dim)
    args = pts + [None] * (3 - dim) + ((frq + [None] * (3 - dim)) if frq else [])
    p.setpts(*args)
    return p, data, dim


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_reference_vectors(oracle, name):
    type_, modes, M, tol, prec = G.CASES[name][:5]
    p, data, dim = _plan(oracle.Plan, name)
    plan = GOLD[name + "/plan"]
    assert [p.ns, p.nc] == [int(plan[0]), int(plan[1])]
    assert p.nf == [int(v) for v in plan[2:2 + dim]]
    if type_ != 3:
        assert np.array_equal(p.perm().astype(np.uint32), GOLD[name + "/perm"])
    got = p.execute(data.reshape(-1))
    err = oracle.relerr(got, GOLD[name + "/out"])
    assert err <= (5e-6 if prec == "f32" else 1e-13), (name, err)
    p.destroy()


def test_golden_file_is_what_the_reference_computes(oracle):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libfinufft_ref.so not built here")
    for name in sorted(G.CASES):
        p, data, dim = _plan(oracle.RefPlan, name)
        out = p.execute(data.reshape(-1))
        assert np.array_equal(out, GOLD[name + "/out"]), name
        if G.CASES[name][0] != 3:
            assert np.array_equal(p.perm().astype(np.uint32), GOLD[name + "/perm"]), name
        p.destroy()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_sort_and_stages_vs_live_reference(oracle, dim, prec):
    """Bigger live comparison: permutation bit-exact on 2e5 points of four distributions, and
    the spread-only / interp-only stages (opts.spreadinterponly, execute.hpp:389-395) against
    the restatement's spread / interp functions."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libfinufft_ref.so not built here")
    from conftest import make_points
    rt = np.float32 if prec == "f32" else np.float64
    ct = np.complex64 if prec == "f32" else np.complex128
    tol = 1e-5 if prec == "f32" else 1e-10
    grid = {1: [4096], 2: [96, 80], 3: [40, 36, 44]}[dim]
    rng = np.random.default_rng(11 + dim)
    M = 200_000
    for kind in ("uniform", "cluster", "wide", "edges"):
        pts = make_points(rng, dim, M, rt, kind, nf=grid)
        for type_ in (1, 2):
            rp = oracle.RefPlan(type_, grid, 1, 1, tol, rt, spread_only=True, nthr=4)
            op = oracle.Plan(type_, grid, 1, 1, tol, rt, spread_only=True, nthr=4)
            rp.setpts(*pts)
            op.setpts(*pts)
            assert np.array_equal(rp.perm(), op.perm()), (kind, type_)
            n_in = M if type_ == 1 else int(np.prod(grid))
            d = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(ct)
            err = oracle.relerr(op.execute(d), rp.execute(d))
            assert err <= (3e-6 if prec == "f32" else 1e-13), (kind, type_, err)
            rp.destroy()
            op.destroy()
