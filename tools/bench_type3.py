"""C5: 3D type 3, single precision, M = N = 1e7, tol 1e-6 (BASELINE.json configs[4]).
Sources uniform in [-pi,pi)^3, target frequencies S_d*(shift_d+u), S_d=107.5, shifts (1.7,-0.5,0.9)
(perftest/perftest.cpp:197-202).  Prints one JSON line: execute ms, setpts ms, points/s."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(3)
pts = [torch.from_numpy(rng.uniform(-np.pi, np.pi, M).astype(np.float32)).cuda() for _ in range(3)]
frq = [torch.from_numpy((107.5 * (sh + rng.uniform(-1, 1, M))).astype(np.float32)).cuda()
       for sh in (1.7, -0.5, 0.9)]
p = F.Plan(3, 3, 1, 1e-6, 1, "complex64", upsampfac=2.0)
p.setpts(*pts, s=frq[0], t=frq[1], u=frq[2])
torch.cuda.synchronize(); t0 = time.perf_counter()
p.setpts(*pts, s=frq[0], t=frq[1], u=frq[2])
torch.cuda.synchronize(); setpts_ms = (time.perf_counter() - t0) * 1e3
c = torch.randn(M, dtype=torch.complex64, device="cuda")
out = p.execute(c)
for _ in range(3):
    p.execute(c, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
K = 5
for _ in range(K):
    p.execute(c, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"workload": f"3D type 3 f32, M=N={M:.3g}, tol=1e-6", "execute_ms": ms,
                  "setpts_ms": setpts_ms, "points_per_s": M / (ms * 1e-3), "plan": p.info()}))
