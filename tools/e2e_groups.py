"""Host-pointer execute (finufft[f]_execute, pinned buffers) for several point-group layouts
(B200_NUFFT_HOST_GROUPS = count, B200_NUFFT_GROUP_FRACS = sizes as fractions of the points).
One process, points generated once.

    python tools/e2e_groups.py --workload c3_t1 --steps 8 \
        --t1 "4:.25,.25,.25,.25;4:.12,.28,.32,.28" --t2 "4:.25,.25,.25,.25"
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import perfdata
from bench import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_t1")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--t1", default="4:.25,.25,.25,.25")
ap.add_argument("--t2", default="")
a = ap.parse_args()
_, modes, M, tol, dtype, ntr = WORKLOADS[a.workload]
rt = np.float32 if dtype == "complex64" else np.float64
rdt = torch.float32 if rt == np.float32 else torch.float64
cdt = torch.complex64 if dtype == "complex64" else torch.complex128
probe = F.Plan(1, modes, 1, tol, 1, dtype, upsampfac=2.0)
nf = probe.info()["nf"]
probe.destroy()
pts_h = perfdata.points(len(modes), M, rt, "uniform", nf)
h_c = torch.empty((M,), dtype=cdt).pin_memory()
h_c.copy_(torch.view_as_complex(torch.randn((M, 2), dtype=rdt)))
h_f = torch.empty(tuple(modes), dtype=cdt).pin_memory()
h_f.copy_(torch.view_as_complex(torch.randn(tuple(modes) + (2,), dtype=rdt)))
for type_, cases in ((1, a.t1), (2, a.t2)):
    for case in [c for c in cases.split(";") if c]:
        g, fr = case.split(":")
        os.environ["B200_NUFFT_HOST_GROUPS"] = g
        os.environ["B200_NUFFT_GROUP_FRACS"] = fr
        hp = F.HostPlan(type_, modes, 1, tol, 1 if type_ == 1 else -1, dtype, upsampfac=2.0,
                        allow_eps_too_small=1)
        hp.setpts(*pts_h[::-1])
        src, dst = (h_c, h_f) if type_ == 1 else (h_f, h_c)
        s_np, d_np = src.numpy(), dst.numpy()
        for _ in range(2):
            hp.execute(s_np, out=d_np)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            hp.execute(s_np, out=d_np)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / a.steps
        print(json.dumps({"workload": a.workload, "type": type_, "groups": int(g), "fracs": fr,
                          "e2e_ms": round(ms, 3), "pts_per_s": M / (ms * 1e-3)}), flush=True)
        hp.destroy()
