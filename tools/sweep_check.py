"""GPU check of the tube-sweep kernels against the generic kernels and the CPU oracle, plus timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F
from oracle import oracle as O


def gpu_exec(type_, ms, M, tol, pts, data, sweep):
    os.environ["B200_NUFFT_SWEEP"] = "1" if sweep else "0"
    p = F.Plan(type_, tuple(ms), 1, tol, 1, "complex64", upsampfac=2.0)
    p.setpts(*pts)
    out = p.execute(data)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    p.enable_profiling(True)
    e0.record(); out = p.execute(data, out); e1.record(); torch.cuda.synchronize()
    st = p.stage_ms()
    res = out.cpu().numpy().reshape(-1)
    p.destroy()
    return res, e0.elapsed_time(e1), st


def check(type_, ms, M, tol, kind="uniform", oracle=True, seed=3):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        pts = [rng.uniform(-np.pi, np.pi, M).astype(np.float32) for _ in range(3)]
    elif kind == "wide":   # far outside [-pi,pi): exercises the fold
        pts = [rng.uniform(-40, 40, M).astype(np.float32) for _ in range(3)]
    elif kind == "cluster":
        pts = [rng.uniform(0, 8 * 2 * np.pi / (2 * m), M).astype(np.float32) for m in ms]
    elif kind == "edge":   # points at and next to the periodic boundary
        pts = [np.where(rng.random(M) < 0.5, np.float32(np.pi), np.float32(-np.pi)).astype(np.float32)
               + rng.choice(np.array([0, 1e-6, -1e-6, 1e-3], dtype=np.float32), M) for _ in range(3)]
    tp = [torch.from_numpy(a).cuda() for a in pts]
    n_in = M if type_ == 1 else int(np.prod(ms))
    data = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
    dd = torch.from_numpy(data if type_ == 1 else data.reshape(ms)).cuda()
    a, ta, sa = gpu_exec(type_, ms, M, tol, tp, dd, True)
    b, tb, sb = gpu_exec(type_, ms, M, tol, tp, dd, False)
    msg = (f"type{type_} ms={ms} M={M} {kind}: sweep {ta:.2f} ms (si {sa['spreadinterp']:.2f}) generic {tb:.2f} ms "
           f"(si {sb['spreadinterp']:.2f}) rel(sweep,generic)={O.relerr(a, b):.2e}")
    if oracle:
        op = O.Plan(type_, list(ms[::-1]), 1, 1, tol, np.float32, sigma=2.0, nthr=O.max_threads())
        op.setpts(*pts[::-1])
        ref = op.execute(data).reshape(-1)
        msg += f" rel(sweep,oracle)={O.relerr(a, ref):.2e} rel(generic,oracle)={O.relerr(b, ref):.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    types = [int(t) for t in os.environ.get("TYPES", "1").split(",")]
    for t in types:
        check(t, (20, 24, 30), 30000, 1e-6)
        check(t, (7, 8, 9), 1000, 1e-6)
        check(t, (16, 16, 16), 5, 1e-6)
        check(t, (32, 32, 32), 200000, 1e-5)      # ns=6
        check(t, (32, 32, 32), 200000, 1e-6, kind="wide")
        check(t, (32, 32, 32), 100000, 1e-6, kind="cluster")
        check(t, (32, 20, 12), 50000, 1e-6, kind="edge")
        check(t, (64, 64, 64), 2000000, 1e-6)
        check(t, (256, 256, 256), 10_000_000, 1e-6, oracle=False)
        check(t, (256, 256, 256), 100_000_000, 1e-6, oracle=False)
