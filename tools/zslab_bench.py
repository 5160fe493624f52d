"""z-slab sharded 3D transform on N GPUs (finufft_b200/zslab.py): parity against the unsharded plan
at a small size, then timing at C3 (256^3 modes, M=1e8 total, tol 1e-6), strong scaling.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
        --master-port P tools/zslab_bench.py [--type 1|2] [--M 1e8] [--steps 5]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import finufft_b200 as F
from finufft_b200.zslab import SlabPlan, slab_of_points

ap = argparse.ArgumentParser()
ap.add_argument("--type", type=int, default=1)
ap.add_argument("--M", type=float, default=1e8)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

# ---------------------------------------------------------------- parity at a small size
modes, M, tol = (64, 48, 40), 200_000, 1e-6
rng = np.random.default_rng(1)
pts = [torch.from_numpy(rng.uniform(-np.pi, np.pi, M).astype(np.float32)).to(dev) for _ in range(3)]
isign = 1 if a.type == 1 else -1
sp = SlabPlan(a.type, modes, tol, isign, "complex64")
own = slab_of_points(pts[0], sp.nf[0], world) == rank
sp.setpts(*[p[own].contiguous() for p in pts])
full = F.Plan(a.type, modes, 1, tol, isign, "complex64", upsampfac=2.0, gpu_device_id=local)
full.setpts(*pts)
if a.type == 1:
    c = torch.view_as_complex(torch.from_numpy(rng.standard_normal((M, 2)).astype(np.float32))).to(dev)
    got = sp.gather_modes(sp.execute(c[own].contiguous()))
    want = full.execute(c)
else:
    fk = torch.view_as_complex(torch.from_numpy(rng.standard_normal(modes + (2,)).astype(np.float32))).to(dev)
    got = sp.execute(fk[:, sp.y_lo:sp.y_hi, :].contiguous())
    want = full.execute(fk)[own]
err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
errs = [None] * world
if world > 1:
    dist.all_gather_object(errs, err)
else:
    errs = [err]
sp.destroy(); full.destroy()
assert max(errs) < 2e-6, f"sharded vs unsharded rel l2 {errs}"

# ---------------------------------------------------------------- timing at C3
modes, Mtot, tol = (256, 256, 256), int(a.M), 1e-6
sp = SlabPlan(a.type, modes, tol, isign, "complex64")
Ml = Mtot // world
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
h = 2 * np.pi / sp.nf[0]
z = (-np.pi + h * (sp.z0 + sp.nz * torch.rand(Ml, device=dev, generator=g))).clamp_(max=float(np.nextafter(np.float32(-np.pi + h * sp.z1), np.float32(-4)))).float()
y = ((torch.rand(Ml, device=dev, generator=g) * 2 - 1) * np.pi).float()
x = ((torch.rand(Ml, device=dev, generator=g) * 2 - 1) * np.pi).float()
assert bool((slab_of_points(z, sp.nf[0], world) == rank).all())
sp.setpts(z, y, x)
if a.type == 1:
    data = torch.view_as_complex(torch.randn((Ml, 2), device=dev, generator=g))
else:
    data = torch.view_as_complex(torch.randn((modes[0], sp.y_hi - sp.y_lo, modes[2], 2), device=dev, generator=g))
for _ in range(a.warmup):
    out = sp.execute(data)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = sp.execute(data)
e1.record()
barrier()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"metric": "NU points/sec (spread+FFT+deconv)", "value": Mtot / (ms * 1e-3), "unit": "points/s",
                      "n_gpus": world, "ms_per_step": ms, "scaling": "strong", "steps": a.steps, "warmup": a.warmup,
                      "config": {"workload": f"3D type {a.type} f32, modes 256^3, M={Mtot:.3g} total, tol=1e-6, "
                                 f"z-slab sharded over {world} GPU(s), points pre-partitioned by slab, "
                                 "outputs left sharded (type 1: y-blocks of modes; type 2: local points)"},
                      "parity_small": {"rel_l2_vs_unsharded": max(errs)}}), flush=True)
sp.destroy()
if world > 1:
    dist.destroy_process_group()
