"""C5 (3D type 3 f32, M = N = 1e7, tol 1e-6, perftest streams) through the host-pointer API with
upsampfac = 2 (fixed) and upsampfac = 0 (sigma3 and the inner sigma chosen at setpts).  Prints one
JSON line per mode: the sigmas / widths / grids chosen, ms per execute (wall, host buffers),
and the error of 8 targets against an f64 direct sum on the GPU."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import finufft_b200 as F
import perfdata

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pts = perfdata.points(3, M, np.float32)
frq = []
for stream, sh in zip("STU", (1.7, -0.5, 0.9)):
    a = np.empty(M, dtype=np.float32)
    perfdata.fill(a, stream, 107.5, sh)
    frq.append(a)
c = perfdata.strengths(M, np.complex64)
out = np.empty(M, dtype=np.complex64)
dp = [torch.from_numpy(p).cuda().double() for p in pts]
df = [torch.from_numpy(f[:8]).cuda().double() for f in frq]
ph = sum(df[d][:, None] * dp[d][None, :] for d in range(3))
want = (torch.polar(torch.ones_like(ph), ph) * torch.from_numpy(c).cuda().to(torch.complex128)[None, :]).sum(1).cpu().numpy()
for sigma in (2.0, 0.0):
    hp = F.HostPlan(3, 3, 1, 1e-6, 1, "complex64", upsampfac=sigma, allow_eps_too_small=1)
    t0 = time.perf_counter()
    hp.setpts(*pts[::-1], **dict(zip("stu", frq[::-1])))
    torch.cuda.synchronize()
    t_set = (time.perf_counter() - t0) * 1e3
    hp.execute(c, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 4
    for _ in range(K):
        hp.execute(c, out=out)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / K
    o, i = hp.info(), hp.info(inner=True)
    err = float(np.linalg.norm(out[:8] - want) / np.linalg.norm(want))
    print(json.dumps({"upsampfac_opt": sigma, "sigma3": o["sigma"], "ns3": o["ns"], "nf": o["nf"],
                      "inner_sigma": i["sigma"], "inner_ns": i["ns"], "inner_nf": i["nf"],
                      "execute_ms_host": round(ms, 3), "setpts_ms_host": round(t_set, 2),
                      "relerr_8_targets": err}), flush=True)
    hp.destroy()
