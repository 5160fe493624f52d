"""C5: 3D type 3, single precision, M = N = 1e7, tol 1e-6 (BASELINE.json configs[4]).
Sources and target frequencies are the reference perftest's streams (perftest/perftest.cpp:197-202:
sources pi*u, targets S_d*(shift_d+u) with S_d = 107.5, shifts (1.7,-0.5,0.9)).  Prints one JSON
line: execute ms, setpts ms, points/s.  Under torchrun (N GPUs) the targets are split across the
ranks (finufft_b200/parallel.py::TargetSplit: no collective), strong scaling, max over ranks."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import finufft_b200 as F
import perfdata
from finufft_b200.parallel import TargetSplit

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def allmax(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


pts = [torch.from_numpy(p).to(dev) for p in perfdata.points(3, M, np.float32)]        # x, y, z
frq = []
for stream, sh in zip("STU", (1.7, -0.5, 0.9)):
    a = np.empty(M, dtype=np.float32)
    perfdata.fill(a, stream, 107.5, sh)
    frq.append(torch.from_numpy(a).to(dev))
ts = TargetSplit(M, lambda: F.Plan(3, 3, 1, 1e-6, 1, "complex64", upsampfac=2.0, gpu_device_id=local))
src, tgt = pts[::-1], frq[::-1]                     # python order: slowest axis first
ts.setpts(src, tgt)
torch.cuda.synchronize(); t0 = time.perf_counter()
ts.setpts(src, tgt)
torch.cuda.synchronize(); setpts_ms = allmax((time.perf_counter() - t0) * 1e3)
c = torch.from_numpy(perfdata.strengths(M, np.complex64)).to(dev)
out = ts.execute_local(c)
for _ in range(3):
    ts.execute_local(c)
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
K = 5
for _ in range(K):
    ts.execute_local(c)
e1.record(); torch.cuda.synchronize()
ms = allmax(e0.elapsed_time(e1) / K)
# accuracy: a few of this rank's targets against an f64 direct sum
lo = ts.lo
sel = torch.arange(0, min(8, ts.hi - ts.lo), device=dev)
ph = torch.zeros((sel.numel(), M), dtype=torch.float64, device=dev)
for d in range(3):
    ph += frq[d][lo + sel].double()[:, None] * pts[d].double()[None, :]
want = (torch.polar(torch.ones_like(ph), ph) * c.to(torch.complex128)[None, :]).sum(1)
err = allmax(float(torch.linalg.norm(out[sel].to(torch.complex128) - want) / torch.linalg.norm(want)))
if rank == 0:
    print(json.dumps({"workload": f"3D type 3 f32, M=N={M:.3g}, tol=1e-6, targets split over {world} GPU(s)",
                      "n_gpus": world, "execute_ms": ms, "setpts_ms": setpts_ms,
                      "points_per_s": M / (ms * 1e-3), "relerr_vs_direct_sum_8_targets": err,
                      "plan": ts.plan.info()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
