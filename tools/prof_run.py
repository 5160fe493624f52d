"""Minimal driver for profiler captures: one plan, setpts, a few executes (no CPU legs)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import perfdata
from bench import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_t1")
ap.add_argument("--M", type=float, default=None)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--dist", default="uniform")
ap.add_argument("--ntr", type=int, default=None)
ap.add_argument("--setpts-ms", action="store_true")
a = ap.parse_args()
type_, modes, M, tol, dtype, ntr = WORKLOADS[a.workload]
M = int(a.M) if a.M else M
ntr = a.ntr or min(ntr, 8)
rt = np.float32 if dtype == "complex64" else np.float64
plan = F.Plan(type_, modes, ntr, tol, 1, dtype, upsampfac=2.0)
nf = plan.info()["nf"]
pts = [torch.from_numpy(p).cuda() for p in perfdata.points(len(modes), M, rt, a.dist, nf)]
plan.setpts(*pts[::-1])
if a.setpts_ms:
    plan.enable_profiling(True)
    plan.setpts(*pts[::-1])
    torch.cuda.synchronize()
    print("setpts_ms", plan.stage_ms()["setpts"], "sort_path", plan.sort_path())
rdt = torch.float32 if rt == np.float32 else torch.float64
shape = ((ntr,) if ntr > 1 else ()) + ((M,) if type_ == 1 else tuple(modes))
data = torch.view_as_complex(torch.randn(shape + (2,), dtype=rdt, device="cuda"))
out = None
for _ in range(a.reps):
    out = plan.execute(data, out)
torch.cuda.synchronize()
print("done", plan.info())
