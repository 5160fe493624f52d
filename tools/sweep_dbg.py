"""Timing experiments for the sweep kernel: B200_SWEEP_DBG bit switches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import finufft_b200 as F
type_ = int(os.environ.get("TYPE", "1"))
for M in (10_000_000, 100_000_000):
    rng = np.random.default_rng(1)
    pts = [torch.from_numpy(rng.uniform(-np.pi, np.pi, M).astype(np.float32)).cuda() for _ in range(3)]
    p = F.Plan(type_, (256, 256, 256), 1, 1e-6, 1, "complex64", upsampfac=2.0)
    p.setpts(*pts)
    p.enable_profiling(True)
    data = torch.randn(M if type_ == 1 else (256, 256, 256), dtype=torch.complex64, device="cuda")
    out = p.execute(data)
    for dbg in [int(x) for x in os.environ.get("DBGS", "0,1,2,4,8,3,15").split(",")]:
        for ns in os.environ.get("NSPLITS", "2").split(","):
            os.environ["B200_SWEEP_DBG"] = str(dbg)
            os.environ["B200_SWEEP_NSPLIT"] = ns
            p.execute(data, out); p.execute(data, out)
            print(f"M={M:.0e} dbg={dbg} nsplit={ns}: spreadinterp {p.stage_ms()['spreadinterp']:.2f} ms", flush=True)
