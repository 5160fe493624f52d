"""The CUDA path (through the C ABI) against the REFERENCE library's own output: the golden
vectors of tests/golden/reference_vectors.npz (made by the reference's CPU code compiled into
oracle/_ref/libfinufft_ref.so, see tests/golden/make_reference_vectors.py) and, when that
library travelled with the snapshot, live runs of it at the BASELINE grid sizes.

Bars: sort permutation bit-exact; outputs within relative l2 of 2 x the requested tolerance
(north_star), in both precisions, types 1, 2 and 3, sigma 2 and 1.25, both mode orders,
ntransf > 1, uniform / clustered / far-out-of-range / boundary points.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_vectors as G  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_gpu_reproduces_reference_vectors(cuda, oracle, name):
    import finufft_b200 as F
    type_, modes, M, tol, prec, ntr, modeord, sigma, kind, N = G.CASES[name]
    ct = np.complex64 if prec == "f32" else np.complex128
    pts, frq, data = G.case_inputs(name)
    dim = len(pts)
    dev = [cuda.from_numpy(p).cuda() for p in pts]
    if type_ == 3:
        gp = F.Plan(3, dim, ntr, tol, 1, ct, upsampfac=sigma)
        gp.setpts(*dev[::-1], **dict(zip("stu", [cuda.from_numpy(f).cuda() for f in frq[::-1]])))
        shape = (ntr, M) if ntr > 1 else (M,)
    else:
        gp = F.Plan(type_, tuple(modes[::-1]), ntr, tol, 1, ct, upsampfac=sigma, modeord=modeord)
        gp.setpts(*dev[::-1])
        inner = (M,) if type_ == 1 else tuple(modes[::-1])
        shape = ((ntr,) if ntr > 1 else ()) + inner
        info = gp.info()
        plan = GOLD[name + "/plan"]
        assert [info["ns"], info["nc"]] == [int(plan[0]), int(plan[1])]
        assert info["nf"] == [int(v) for v in plan[2:2 + dim]]
        assert np.array_equal(gp.sort_permutation(), GOLD[name + "/perm"])
    got = gp.execute(cuda.from_numpy(data.reshape(shape)).cuda()).cpu().numpy()
    err = oracle.relerr(got, GOLD[name + "/out"])
    print(f"\n[{name}] gpu-vs-reference {err:.3e} (tol {tol:g})")
    assert err <= 2 * tol, (name, err)
    gp.destroy()


@pytest.mark.parametrize("type_", [1, 2])
def test_c3_grid_vs_live_reference(cuda, oracle, type_):
    """BASELINE configs[2] grid (256^3 modes, fine grid 512^3, ns = 7, tol 1e-6, f32) with
    M = 4e6 perftest-generated points: GPU vs the reference library itself, plus the float floor
    (reference f32 vs reference f64) the north_star bar has to be read against."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libfinufft_ref.so not in this snapshot")
    import finufft_b200 as F
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import perfdata
    tol, M, modes = 1e-6, 4_000_000, (256, 256, 256)
    pts = perfdata.points(3, M, np.float32)
    n_in = M if type_ == 1 else 256 ** 3
    data = perfdata.strengths(n_in, np.complex64, "C" if type_ == 1 else "FK")
    nthr = oracle.max_threads()
    rp = oracle.RefPlan(type_, list(modes), 1, 1, tol, np.float32, nthr=nthr)
    rp.setpts(*pts)
    want = rp.execute(data)
    rp64 = oracle.RefPlan(type_, list(modes), 1, 1, tol, np.float64, nthr=nthr)
    rp64.setpts(*[p.astype(np.float64) for p in pts])
    want64 = rp64.execute(data.astype(np.complex128))
    gp = F.Plan(type_, modes, 1, tol, 1, np.complex64, upsampfac=2.0)
    gp.setpts(*[cuda.from_numpy(p).cuda() for p in pts[::-1]])
    assert np.array_equal(gp.sort_permutation().astype(np.int64), rp.perm())
    shape = (M,) if type_ == 1 else modes
    got = gp.execute(cuda.from_numpy(data.reshape(shape)).cuda()).cpu().numpy()
    err, floor = oracle.relerr(got, want), oracle.relerr(want, want64)
    err64 = oracle.relerr(got, want64)
    print(f"\n[C3 grid type {type_}] gpu-vs-reference-f32 {err:.3e}  reference f32-vs-f64 "
          f"{floor:.3e}  gpu-vs-reference-f64 {err64:.3e}")
    assert err <= max(2 * tol, 3 * floor)
    assert err64 <= 1.5 * floor + 2 * tol   # the GPU is as close to the truth as the reference
    for p in (rp, rp64, gp):
        p.destroy()
