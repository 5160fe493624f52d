"""Parity of the sharded (z-slab) plan against the unsharded plan on the same GPU type.

    python tools/sharded_check.py                       # world = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29611 tools/sharded_check.py      # world = N

Every rank builds the same full point set, keeps every world-th chunk of it, and the sharded
result is compared with the ordinary single-GPU plan run on the full set.  Exit code 0 = all
cases within 2*tol; one line per case on rank 0.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finufft_b200  # noqa: E402
from finufft_b200.sharded import ShardedPlan  # noqa: E402


def relerr(a, b):
    return float(torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1)))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    big = "--big" in sys.argv
    cases = []
    for dtype, tol in (("complex64", 1e-6), ("complex128", 1e-9)):
        for typ in (1, 2):
            for dist_name in ("uniform", "cluster", "slabbed"):
                cases.append((dtype, tol, typ, dist_name, (40, 36, 48), 60000, 0))
    cases.append(("complex64", 1e-6, 1, "uniform", (64, 64, 64), 400000, 1))   # FFT order
    cases.append(("complex64", 1e-4, 2, "uniform", (33, 47, 50), 50000, 0))    # odd sizes, ns=5
    if big:
        cases.append(("complex64", 1e-6, 1, "uniform", (256, 256, 256), 4000000, 0))
        cases.append(("complex64", 1e-6, 2, "uniform", (256, 256, 256), 4000000, 0))
        cases.append(("complex64", 1e-6, 1, "cluster", (256, 256, 256), 4000000, 0))
    worst = 0.0
    failed = 0
    for dtype, tol, typ, dname, modes, M, modeord in cases:
        rt = torch.float32 if dtype == "complex64" else torch.float64
        ct = torch.complex64 if dtype == "complex64" else torch.complex128
        g = torch.Generator(device="cpu").manual_seed(1234 + M + typ)
        pts = (torch.rand((3, M), generator=g, dtype=torch.float64) * 2 - 1) * np.pi
        sp = ShardedPlan(typ, modes, eps=tol, dtype=dtype, modeord=modeord,
                         group=None)
        info = sp.info()
        nf3 = info["nf"][2]
        routed = False
        if dname == "cluster":      # every point inside 8 fine-grid cells per dimension
            h = 2 * np.pi / np.array(info["nf"][::-1], dtype=np.float64)
            pts = (pts / np.pi + 1) * 4 * torch.from_numpy(h)[:, None] + 0.3
        if dname == "slabbed":      # caller pre-partitions: rank r gets the points of its slab
            # points within 1e-3 of a plane boundary are dropped so that the float fold of the
            # library and this float64 one agree on every owner
            zz = pts[0] / (2 * np.pi) + 0.5
            zf = (zz - torch.floor(zz)) * nf3
            keep = (zf - torch.floor(zf) > 1e-3) & (zf - torch.floor(zf) < 1 - 1e-3)
            pts, zf = pts[:, keep], zf[keep]
            M = pts.shape[1]
            plane = torch.clamp(zf.long(), max=nf3 - 1)
            mine = (plane >= info["z0"]) & (plane < info["z0"] + info["nz"])
            idx = torch.nonzero(mine).reshape(-1)
            routed = True
        pts = pts.to(rt)
        if routed:
            pass
        else:
            idx = torch.arange(rank, M, world)
        zl, yl, xl = (pts[d][idx].to(dev) for d in range(3))
        c_full = (torch.randn(M, generator=g, dtype=torch.float64)
                  + 1j * torch.randn(M, generator=g, dtype=torch.float64)).to(ct)
        f_full = (torch.randn(modes, generator=g, dtype=torch.float64)
                  + 1j * torch.randn(modes, generator=g, dtype=torch.float64)).to(ct)
        ref = finufft_b200.Plan(typ, modes, 1, tol, dtype=dtype, modeord=modeord,
                                gpu_device_id=local)
        ref.setpts(*(pts[d].to(dev) for d in range(3)))
        sp.setpts(zl, yl, xl, routed=routed)
        info = sp.info()
        if typ == 1:
            want = ref.execute(c_full.to(dev))
            blk = sp.execute(c_full[idx].to(dev))
            got = sp.gather_modes(blk)
            err = relerr(got, want)
            # the block alone must equal the slice of the full array
            err = max(err, relerr(blk, want[:, info["ylo"]:info["yhi"], :]) if blk.numel() else 0)
        else:
            want = ref.execute(f_full.to(dev))
            blk = sp.slice_modes(f_full.to(dev))
            assert torch.equal(blk, f_full.to(dev)[:, info["ylo"]:info["yhi"], :])
            got = sp.execute(blk)
            wl = want[idx.to(dev)]
            num = torch.linalg.norm(got - wl) ** 2
            den = torch.linalg.norm(wl) ** 2
            if world > 1:
                both = torch.stack([num, den])
                dist.all_reduce(both)
                num, den = both[0], both[1]
            err = float(torch.sqrt(num / den))
        if world > 1:
            e = torch.tensor([err], device=dev)
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            err = float(e)
        ok = err <= 2 * tol
        failed += 0 if ok else 1
        worst = max(worst, err / tol)
        if rank == 0:
            st = sp.stage_ms()
            print(f"[sharded W={world}] {dtype} type {typ} {dname:8s} modes {modes} M={M} "
                  f"modeord={modeord} mode={info['mode']} win={info['win_org']}+{info['win_n']} "
                  f"M_local={info['M_local']} relerr {err:.3e} {'ok' if ok else 'FAIL'} "
                  f"exec {st['execute']:.3f} ms", flush=True)
        sp.destroy()
        ref.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(f"sharded_check: {len(cases) - failed}/{len(cases)} ok, worst err/tol {worst:.3f}")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
