"""Ad-hoc GPU sanity script (not collected by pytest)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F
from oracle import oracle as O

def run(type_, dim, dt, tol, ms, M, ntr=1, modeord=0, kind="uniform", **kw):
    rng = np.random.default_rng(7)
    rt = np.float32 if dt == "f" else np.float64
    ct = np.complex64 if dt == "f" else np.complex128
    pts = [rng.uniform(-np.pi, np.pi, M).astype(rt) for _ in range(dim)]
    p = F.Plan(type_, tuple(ms), ntr, tol, 1, ct, modeord=modeord, upsampfac=2.0, **kw)
    tp = [torch.from_numpy(a).cuda() for a in pts]
    p.setpts(*tp)
    info = p.info()
    # oracle uses library axis order: x = last python axis
    op = O.Plan(type_, list(ms[::-1]), 1, ntr, tol, rt, sigma=2.0, modeord=modeord, nthr=8)
    lp = pts[::-1] + [None] * (3 - dim)
    op.setpts(*lp)
    perm_ok = np.array_equal(p.sort_permutation().astype(np.int64), op.perm())
    N = int(np.prod(ms))
    if type_ == 1:
        c = (rng.standard_normal((ntr, M)) + 1j * rng.standard_normal((ntr, M))).astype(ct)
        cin = c if ntr > 1 else c[0]
        out = p.execute(torch.from_numpy(cin).cuda()).cpu().numpy()
        ref = op.execute(c)
    else:
        fk = (rng.standard_normal((ntr,) + tuple(ms)) + 1j * rng.standard_normal((ntr,) + tuple(ms))).astype(ct)
        fin = fk if ntr > 1 else fk[0]
        out = p.execute(torch.from_numpy(fin).cuda()).cpu().numpy()
        ref = op.execute(fk)
    err = O.relerr(out.reshape(-1), ref.reshape(-1))
    print(f"type{type_} dim{dim} {dt} tol={tol} ms={ms} M={M} ntr={ntr} ns={info['ns']} nf={info['nf']} "
          f"nsub={info['nsub']} perm_ok={perm_ok} relerr_vs_oracle={err:.3e}", flush=True)
    return err

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for dt, tol in (("f", 1e-5), ("d", 1e-9)):
        for type_ in (1, 2):
            run(type_, 1, dt, tol, (200,), 5000)
            run(type_, 2, dt, tol, (60, 44), 20000)
            run(type_, 3, dt, tol, (20, 24, 30), 30000)
    run(1, 3, "f", 1e-6, (32, 32, 32), 200000, ntr=2)
    run(2, 3, "f", 1e-6, (32, 32, 32), 200000, ntr=2, modeord=1)
    # timing at moderate scale
    for type_ in (1, 2):
        M = 10_000_000
        rng = np.random.default_rng(1)
        pts = [torch.from_numpy(rng.uniform(-np.pi, np.pi, M).astype(np.float32)).cuda() for _ in range(3)]
        p = F.Plan(type_, (256, 256, 256), 1, 1e-6, 1, "complex64", upsampfac=2.0)
        torch.cuda.synchronize(); t0 = time.time()
        p.setpts(*pts); torch.cuda.synchronize(); t1 = time.time()
        if type_ == 1:
            data = torch.randn(M, dtype=torch.complex64, device="cuda")
        else:
            data = torch.randn((256, 256, 256), dtype=torch.complex64, device="cuda")
        out = p.execute(data); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = p.execute(data, out); e1.record(); torch.cuda.synchronize()
        print(f"type{type_} 3D 256^3 M=1e7: setpts {1e3*(t1-t0):.1f} ms, execute {e0.elapsed_time(e1):.2f} ms", flush=True)
