"""CPU-only, world_size 2 over gloo: the z-slab sharding of one 3D transform
(tests/zslab_model.py, the host-side model of csrc/slab.cu) — slab ownership, ghost-plane ring exchange, slab->pencil all_to_all,
mode selection and deconvolution — with the CPU oracle standing in for the GPU spreader /
interpolator.  The sharded result must equal the oracle's unsharded transform."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

MODES = (16, 12, 10)   # python order (ms3, ms2, ms1)
TOL = 1e-9


def test_mode_indices_and_bounds():
    from zslab_model import mode_indices, slab_bounds, split_even
    ix, ak = mode_indices(20, 7)            # k = -3..3
    assert ix.tolist() == [17, 18, 19, 0, 1, 2, 3] and ak.tolist() == [3, 2, 1, 0, 1, 2, 3]
    ix, ak = mode_indices(16, 8)            # k = -4..3
    assert ix.tolist() == [12, 13, 14, 15, 0, 1, 2, 3] and ak.tolist() == [4, 3, 2, 1, 0, 1, 2, 3]
    assert slab_bounds(512, 8, 3) == (192, 256)
    with pytest.raises(ValueError):
        slab_bounds(36, 2, 0)               # 18 planes per slab is not a whole number of bins
    assert split_even(10, 4) == ([0, 3, 6, 8], [3, 3, 2, 2])


class _OracleLocal:
    """Spread-only / interp-only operator on a periodic grid, CPU, torch tensors in and out."""

    def __init__(self, type_, shape, isign):
        from oracle import oracle as O
        self.type, self.shape = type_, tuple(shape)
        self.p = O.Plan(type_, list(shape[::-1]), isign, 1, TOL, np.float64, sigma=2.0,
                        spread_only=True, nthr=1)

    def setpts(self, z, y, x):
        self.p.setpts(x.numpy(), y.numpy(), z.numpy())

    def execute(self, data, out=None):
        res = self.p.execute(np.ascontiguousarray(data.numpy()).reshape(-1))
        res = torch.from_numpy(np.asarray(res).reshape(self.shape if self.type == 1 else (-1,)))
        if out is not None:
            out.copy_(res)
            return out
        return res


def _points(M):
    rng = np.random.default_rng(5)
    return [torch.from_numpy(rng.uniform(-np.pi, np.pi, M)) for _ in range(3)]  # z, y, x


def _worker(rank, world, port, type_, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import build as ob
    ob.build_oracle()
    from oracle import oracle as O
    from zslab_model import SlabPlan, route_points, slab_of_points
    M = 3000
    z, y, x = _points(M)
    rng = np.random.default_rng(6)
    isign = 1 if type_ == 1 else -1
    sp = SlabPlan(type_, MODES, TOL, isign, "complex128", device=torch.device("cpu"),
                  make_local=lambda shape: _OracleLocal(type_, shape, isign))
    ref = O.Plan(type_, list(MODES[::-1]), isign, 1, TOL, np.float64, sigma=2.0, nthr=1)
    ref.setpts(x.numpy(), y.numpy(), z.numpy())
    owner = slab_of_points(z, sp.nf[0], world)
    if type_ == 1:
        c = torch.from_numpy(rng.standard_normal(M) + 1j * rng.standard_normal(M))
        # every rank starts with an arbitrary half of the points and routes them to their owners
        mine = torch.arange(M) % world == rank
        lz, ly, lx, lc = route_points(z[mine], y[mine], x[mine], c[mine], sp.nf[0])
        assert bool((slab_of_points(lz, sp.nf[0], world) == rank).all())
        sp.setpts(lz, ly, lx)
        got = sp.gather_modes(sp.execute(lc)).numpy()
        want = np.asarray(ref.execute(c.numpy())).reshape(MODES)
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    else:
        fk = rng.standard_normal(MODES) + 1j * rng.standard_normal(MODES)
        sel = owner == rank
        sp.setpts(z[sel], y[sel], x[sel])
        got = sp.execute(torch.from_numpy(fk[:, sp.y_lo:sp.y_hi, :].copy())).numpy()
        want = np.asarray(ref.execute(fk.reshape(-1))).reshape(-1)[sel.numpy()]
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    q.put((rank, err, int(sp.z0), int(sp.z1)))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("type_", [1, 2])
def test_zslab_matches_unsharded_oracle(type_):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, type_, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, z0, z1 in sorted(res):
        assert (z0, z1) == (rank * 16, rank * 16 + 16)
        assert err < 1e-8, f"rank {rank}: rel l2 error {err:.2e} vs the unsharded oracle"
