"""CPU-only, world_size 2 over gloo: the N>1 host logic (finufft_b200/parallel.py): vectors of
a batched transform are split across ranks with no data-path collective; one all_gather only
when every rank wants the full result."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_split_transforms():
    from finufft_b200.parallel import split_transforms
    assert split_transforms(64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    assert split_transforms(5, 2) == [(0, 3), (3, 5)]
    assert split_transforms(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert split_transforms(0, 3) == [(0, 0)] * 3
    for ntr in range(0, 40):
        for w in range(1, 9):
            s = split_transforms(ntr, w)
            assert s[0][0] == 0 and s[-1][1] == ntr
            assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
            sizes = [hi - lo for lo, hi in s]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        split_transforms(3, 0)


class _FakePlan:
    """Stands in for a GPU plan: a deterministic linear map per vector."""

    def __init__(self, n_local):
        self.n_local = n_local
        self.w = None

    def setpts(self, x):
        self.w = torch.outer(torch.arange(1, 5, dtype=torch.float64), x.double()).to(torch.complex128)

    def execute(self, data):
        return data @ self.w.T if data.dim() == 2 else self.w @ data


def _worker(rank, world, port, ntr, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from finufft_b200.parallel import BatchSplit
    torch.manual_seed(0)
    x = torch.linspace(-1, 1, 7)
    data = torch.randn(ntr, 7, dtype=torch.complex128)
    bs = BatchSplit(ntr, lambda n: _FakePlan(n))
    bs.setpts(x)
    full = bs.execute_gathered(data, (4,))
    ref = _FakePlan(ntr)
    ref.setpts(x)
    want = ref.execute(data)
    ok = torch.allclose(full, want) and tuple(full.shape) == (ntr, 4)
    mine = bs.execute_local(data)
    lo, hi = bs.slices[rank]
    ok = ok and ((mine is None and hi == lo) or torch.allclose(mine, want[lo:hi]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("ntr", [5, 1, 8])
def test_batch_split_two_ranks_gloo(ntr):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ntr, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


class _OracleType3:
    """CPU stand-in with the interface of a type-3 finufft_b200.Plan: the oracle's own type 3."""

    def __init__(self, dim, tol):
        from oracle import oracle as O
        self.O, self.dim, self.tol, self.p = O, dim, tol, None

    def setpts(self, *src, s=None, t=None, u=None):
        import numpy as np
        O = self.O
        self.p = O.Plan(3, [1] * self.dim, 1, 1, self.tol, np.float64, nthr=1, dim=self.dim)
        frq = [a for a in (s, t, u) if a is not None]
        # python order (slowest first) -> library order, like finufft_b200.Plan.setpts
        pts = [a.numpy() for a in src][::-1] + [None] * (3 - self.dim)
        fr = [a.numpy() for a in frq][::-1] + [None] * (3 - self.dim)
        self.p.setpts(*pts, *fr)

    def execute(self, c):
        return torch.from_numpy(self.p.execute(c.numpy()))


def _worker_t3(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from finufft_b200.parallel import TargetSplit
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    M, N, tol = 3000, 2501, 1e-9
    src = [torch.from_numpy(rng.uniform(-np.pi, np.pi, M)) for _ in range(2)]
    tgt = [torch.from_numpy(25.0 * rng.uniform(-1, 1, N) + sh) for sh in (3.0, -7.0)]
    c = torch.from_numpy(rng.standard_normal(M) + 1j * rng.standard_normal(M))
    ts = TargetSplit(N, lambda: _OracleType3(2, tol))
    ts.setpts(src, tgt)
    full = ts.execute_gathered(c)
    want = O.dirft(3, src[1].numpy(), src[0].numpy(), None, c.numpy(), 1,
                   s=tgt[1].numpy(), t=tgt[0].numpy())
    err = float(np.linalg.norm(full.numpy() - want) / np.linalg.norm(want))
    q.put((rank, bool(err < 1e-8 and full.shape[0] == N)))
    dist.destroy_process_group()


def test_type3_target_split_two_ranks_gloo():
    """Type 3 sharded by targets (finufft_b200/parallel.py::TargetSplit), world size 2 over gloo,
    the oracle's type 3 as the local plan, against the reference's direct sum."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_t3, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
