"""Multi-GPU sharding of the hot path across one 8xB200 box (one process per GPU).

The path shards where the reference's own batching does: the `ntransf` strength vectors of a
vectorised transform are independent given the points (include/finufft/execute.hpp:376-382:
cjb = cj + b*nj, fkb = fk + b*N).  Each rank keeps a full copy of the points (setpts is
replicated; 20 B/point once) and executes its contiguous slice of the vectors on its own GPU
with no data-path collective.  Only if the caller wants every rank to hold the full result is
one all_gather issued at the end (NCCL over NVLink on GPUs, gloo in the CPU tests).

The single-large-transform case (z-slab decomposition of the fine grid with ghost-plane
exchange and a distributed FFT) is `finufft_b200.sharded.ShardedPlan`, which
binds the C++ / NCCL implementation behind include/b200_sharded.h.

`bench.py --workload c4_t1 --gpus N` runs this split on the BASELINE batched config (2D f64,
512^2 modes, M = 1e7, ntransf = 64).
"""
from typing import Callable, List, Tuple


def split_transforms(ntr: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) slice of the ntr vectors for every rank; sizes differ by at
    most one, earlier ranks get the larger share, empty slices allowed when world > ntr."""
    if ntr < 0 or world < 1:
        raise ValueError("ntr >= 0 and world >= 1 required")
    base, extra = divmod(ntr, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


class BatchSplit:
    """Runs a vectorised transform with its vectors split across the ranks of a process group.

    make_plan(n_local) -> object with setpts(*pts) and execute(data) (a finufft_b200.Plan on
    this rank's GPU).  data_in is the FULL (ntr, ...) input, present on every rank (or only the
    local slice if `local_only`).
    """

    def __init__(self, ntr: int, make_plan: Callable, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ntr = ntr
        self.slices = split_transforms(ntr, self.world)
        self.lo, self.hi = self.slices[self.rank]
        self.plan = make_plan(self.hi - self.lo) if self.hi > self.lo else None

    def setpts(self, *pts):
        if self.plan is not None:
            self.plan.setpts(*pts)

    def execute_local(self, data_full):
        """This rank's slice of the outputs (None if the slice is empty)."""
        if self.plan is None:
            return None
        local = data_full[self.lo:self.hi]
        if self.hi - self.lo == 1:
            return self.plan.execute(local[0]).unsqueeze(0)
        return self.plan.execute(local.contiguous())

    def execute_gathered(self, data_full, out_shape_per_vector):
        """Full (ntr, ...) result on every rank: local execute + one all_gather."""
        import torch
        mine = self.execute_local(data_full)
        if self.world == 1:
            return mine
        nmax = max(hi - lo for lo, hi in self.slices)
        ref = mine if mine is not None else data_full
        buf = torch.zeros((nmax,) + tuple(out_shape_per_vector), dtype=data_full.dtype,
                          device=ref.device)
        if mine is not None:
            buf[: mine.shape[0]] = mine
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(parts, buf, group=self.group)
        return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, self.slices)], dim=0)


class TargetSplit:
    """Type 3 (nonuniform -> nonuniform) over the ranks of a process group.

    The outputs of a type-3 transform are independent given the sources: f_k = sum_j c_j
    exp(+-i s_k . x_j) (include/finufft/execute.hpp:432-558), so the N target frequencies are split
    into contiguous slices, every rank keeps all M sources and evaluates its slice with an
    ordinary type-3 plan on its own GPU: no data-path collective, outputs stay sharded (or one
    all_gather when every rank wants all of them).  Each rank's plan chooses its own spreading
    grid from the extent of ITS targets (setpts.hpp:163-319), which only shrinks the work.
    (The finer decomposition of SURVEY.md 8(e) - outer spread on z-slabs, a slab -> block
    transpose of the spreading grid, a sharded inner type 2 - is not built.)

    make_plan() -> object with setpts(*src, s=.., t=.., u=..) and execute(c) (a type-3
    finufft_b200.Plan on this rank's GPU).
    """

    def __init__(self, n_targets: int, make_plan: Callable, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_targets = n_targets
        self.slices = split_transforms(n_targets, self.world)
        self.lo, self.hi = self.slices[self.rank]
        self.plan = make_plan() if self.hi > self.lo else None

    def setpts(self, sources, targets):
        """sources, targets: sequences of coordinate arrays in the plan's own argument order;
        the targets are the FULL arrays, this rank keeps [lo, hi)."""
        if self.plan is not None:
            names = ("s", "t", "u")[: len(targets)]
            self.plan.setpts(*sources, **{n: a[self.lo:self.hi].contiguous()
                                          for n, a in zip(names, targets)})

    def execute_local(self, c):
        return None if self.plan is None else self.plan.execute(c)

    def execute_gathered(self, c):
        import torch
        mine = self.execute_local(c)
        if self.world == 1:
            return mine
        nmax = max(hi - lo for lo, hi in self.slices)
        buf = torch.zeros((nmax,), dtype=c.dtype, device=c.device)
        if mine is not None:
            buf[: mine.shape[0]] = mine
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(parts, buf, group=self.group)
        return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, self.slices)], dim=0)
