"""TEST INFRASTRUCTURE: host-side model of the z-slab decomposition, for the CPU tests only
(tests/test_zslab_gloo_cpu.py); not part of the finufft_b200 package.

The product path is csrc/slab.cu behind include/b200_sharded.h (python: finufft_b200.sharded.
ShardedPlan): C++ driving the library's own kernels, cuFFT and NCCL.  This module restates the
same decomposition with torch.distributed collectives and an injectable local spreader, so that
tests/test_zslab_gloo_cpu.py can check the exchange logic (ghost planes, slab<->pencil
transpose, mode blocks, deconvolution factors) on CPU with the `gloo` backend and the oracle
as the local spreader.  Nothing in bench.py or the GPU tests uses it.

One large 3D type-1 / type-2 transform sharded across the GPUs of one box by z-slabs of the
fine grid.  One process per GPU; `torch.distributed` (NCCL over NVLink on
GPUs, gloo in the CPU tests) carries the two exchange steps the path really has:

  type 1   spread local points into the local copy of the fine grid (this rank's planes plus
           ns/2 ghost planes below and ns-ns/2 above, periodic)
           -> ghost planes to the ring neighbours, which add them            [send/recv]
           -> 2D FFT (x,y) of the owned planes, keep the ms1 x ms2 modes wanted
           -> slab -> pencil transpose                                        [all_to_all]
           -> 1D FFT along z, keep ms3 modes, deconvolve by phihat1*phihat2*phihat3
           => this rank's y-range of the mode array, fk[:, y_lo:y_hi, :]
  type 2   the mirror image, ending with interpolation at the local points.

Grid geometry, mode ordering and the deconvolution factors are the single-GPU path's
(include/finufft/execute.hpp:69-237; phihat from makeplan.hpp:39-108, which already carries the
(-1)^k of the half-period grid shift), so the result equals the unsharded transform up to FFT
rounding.  The spreader / interpolator is a `finufft_b200.Plan` in `gpu_spreadinterponly` mode on
a grid of the fine-grid size (include/finufft/execute.hpp:389-395 semantics); tests inject a CPU
stand-in with the same interface.  The FFTs are library calls (cuFFT through torch.fft).

Points must already live on the rank that owns their slab: rank r owns fine-grid planes
[r*nz, (r+1)*nz), nz = nf3 / world (a multiple of the 4-plane bin depth), i.e. the points with
fold(z) in that range (`slab_of_points` computes the owner; `route_points` does the exchange
for callers whose points are not yet partitioned).
"""
import math

import numpy as np


def mode_indices(nf, ms):
    """Fine-grid FFT index of every wanted mode k = -(ms//2) .. (ms-1)//2, in increasing k
    (modeord 0), and |k| for the phihat lookup (execute.hpp:98-118)."""
    k = np.arange(ms) - ms // 2
    return (k % nf).astype(np.int64), np.abs(k).astype(np.int64)


def slab_bounds(nf3, world, rank, depth=4):
    nz = nf3 // world
    if nz * world != nf3 or nz % depth:
        raise ValueError(f"fine grid depth {nf3} must split into {world} slabs of whole bins")
    return rank * nz, (rank + 1) * nz


def split_even(n, world):
    base, extra = divmod(n, world)
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    starts = [sum(sizes[:r]) for r in range(world)]
    return starts, sizes


def slab_of_points(z, nf3, world):
    """Owner rank of every point: fold z to [0, nf3) exactly as the library does
    (include/finufft/simd.hpp:318-325) and divide by the slab depth."""
    import torch
    r = torch.addcmul(torch.full_like(z, 0.5), z, torch.full_like(z, 0.15915494309189535))
    zz = (r - torch.floor(r)) * nf3
    return torch.clamp((zz / (nf3 // world)).to(torch.int64), max=world - 1)


class SlabPlan:
    """Distributed 3D plan.  n_modes = (ms3, ms2, ms1) in python (C) order like finufft_b200.Plan.

    make_local(grid_shape) -> object with setpts(z, y, x) and execute(data, out=None) acting as a
    spread-only (type 1) or interp-only (type 2) operator on a periodic grid of `grid_shape`
    (python order).  Default: finufft_b200.Plan(..., gpu_spreadinterponly=1) on the current GPU.
    """

    def __init__(self, nufft_type, n_modes, eps, isign=None, dtype="complex64", group=None,
                 make_local=None, device=None, upsampfac=2.0):
        import torch
        import torch.distributed as dist
        from finufft_b200 import hostmath
        if nufft_type not in (1, 2) or len(n_modes) != 3:
            raise ValueError("SlabPlan handles 3D transforms of type 1 and 2")
        self.torch, self.dist, self.group = torch, dist, group
        self.type = nufft_type
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ms = tuple(int(m) for m in n_modes)
        self.isign = (1 if nufft_type == 1 else -1) if isign is None else (1 if isign >= 0 else -1)
        self.cdtype = torch.complex64 if np.dtype(dtype) == np.complex64 else torch.complex128
        rt = np.float32 if self.cdtype == torch.complex64 else np.float64
        err, self.ns, _, table = hostmath.kernel(eps, 3, nufft_type, upsampfac, rt)
        if err:
            raise RuntimeError(f"kernel selection failed with code {err}")
        self.nf = tuple(hostmath.fine_grid(upsampfac, m, self.ns) for m in self.ms)
        self.z0, self.z1 = slab_bounds(self.nf[0], self.world, self.rank)
        self.nz = self.z1 - self.z0
        self.below, self.above = self.ns // 2, self.ns - self.ns // 2
        if self.world > 1 and self.nz < max(self.below, self.above):
            raise ValueError("slabs thinner than the ghost depth")
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
            else torch.device("cpu"))
        self.ystart, self.ysize = split_even(self.ms[1], self.world)
        self.y_lo = self.ystart[self.rank]
        self.y_hi = self.y_lo + self.ysize[self.rank]
        dev = self.device
        self.idx, self.phi = [], []
        for d in range(3):
            ix, ak = mode_indices(self.nf[d], self.ms[d])
            ph = hostmath.fseries(self.nf[d], table)
            self.idx.append(torch.from_numpy(ix).to(dev))
            self.phi.append(torch.from_numpy(ph[ak].astype(rt)).to(dev))
        # deconvolution factor of this rank's pencil block (ms3, my_y, ms1)
        self.dec = 1.0 / (self.phi[0][:, None, None] * self.phi[1][None, self.y_lo:self.y_hi, None]
                          * self.phi[2][None, None, :])
        if make_local is None:
            from finufft_b200.plan import Plan

            def make_local(shape):
                return Plan(nufft_type, shape, 1, eps, self.isign,
                            "complex64" if self.cdtype == torch.complex64 else "complex128",
                            upsampfac=upsampfac, gpu_spreadinterponly=1,
                            gpu_device_id=dev.index or 0)
        self.local = make_local(self.nf)
        self.grid = torch.zeros(self.nf, dtype=self.cdtype, device=dev)

    # ------------------------------------------------------------------ points
    def setpts(self, z, y, x):
        """Local points (python order: slowest axis first).  Every z must fold into this rank's
        slab."""
        self.M = z.numel()
        self.local.setpts(z, y, x)

    # ------------------------------------------------------------------ exchanges
    def _ring(self, send_up, send_down):
        """send_up -> rank+1, send_down -> rank-1; returns (from rank-1, from rank+1)."""
        torch, dist = self.torch, self.dist
        nxt, prv = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        from_prev, from_next = torch.empty_like(send_up), torch.empty_like(send_down)
        ops = [dist.P2POp(dist.isend, send_up, nxt, self.group),
               dist.P2POp(dist.irecv, from_prev, prv, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ops = [dist.P2POp(dist.isend, send_down, prv, self.group),
               dist.P2POp(dist.irecv, from_next, nxt, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return from_prev, from_next

    def _planes(self, lo, n):
        torch = self.torch
        return (torch.arange(lo, lo + n, device=self.device) % self.nf[0])

    def _fft(self, t, dims):
        fft = self.torch.fft
        if self.isign > 0:  # unnormalised e^{+i...}
            return fft.ifftn(t, dim=dims, norm="forward")
        return fft.fftn(t, dim=dims, norm="backward")

    # ------------------------------------------------------------------ type 1
    def _execute1(self, c):
        torch, dist = self.torch, self.dist
        W, nz, (ms3, ms2, ms1) = self.world, self.nz, self.ms
        grid = self.local.execute(c, out=self.grid)  # zeroes, then spreads
        if W > 1:
            up = grid[self._planes(self.z1, self.above)].contiguous()
            down = grid[self._planes(self.z0 - self.below, self.below)].contiguous()
            from_prev, from_next = self._ring(up, down)
            grid[self.z0:self.z0 + self.above] += from_prev
            grid[self.z1 - self.below:self.z1] += from_next
        slab = self._fft(grid[self.z0:self.z1], (1, 2))
        slab = slab[:, self.idx[1]][:, :, self.idx[2]]              # (nz, ms2, ms1)
        if W > 1:
            send = slab.permute(1, 0, 2).contiguous()               # (ms2, nz, ms1)
            my = self.ysize[self.rank]
            recv = torch.empty((W * my, nz, ms1), dtype=self.cdtype, device=self.device)
            dist.all_to_all_single(recv, send, output_split_sizes=[my] * W,
                                   input_split_sizes=self.ysize, group=self.group)
            pencil = recv.view(W, my, nz, ms1).permute(0, 2, 1, 3).reshape(W * nz, my, ms1)
        else:
            pencil = slab
        pencil = self._fft(pencil, (0,))[self.idx[0]]               # (ms3, my, ms1)
        return pencil * self.dec

    # ------------------------------------------------------------------ type 2
    def _execute2(self, fk_local):
        torch, dist = self.torch, self.dist
        W, nz, (ms3, ms2, ms1) = self.world, self.nz, self.ms
        nf3, nf2, nf1 = self.nf
        my = self.ysize[self.rank]
        pencil = torch.zeros((nf3, my, ms1), dtype=self.cdtype, device=self.device)
        pencil[self.idx[0]] = fk_local * self.dec
        pencil = self._fft(pencil, (0,))                            # (nf3, my, ms1)
        if W > 1:
            send = pencil.view(W, nz, my, ms1).permute(0, 2, 1, 3).contiguous().view(W * my, nz, ms1)
            recv = torch.empty((ms2, nz, ms1), dtype=self.cdtype, device=self.device)
            dist.all_to_all_single(recv, send, output_split_sizes=self.ysize,
                                   input_split_sizes=[my] * W, group=self.group)
            slab_modes = recv.permute(1, 0, 2)                      # (nz, ms2, ms1)
        else:
            slab_modes = pencil
        slab = torch.zeros((nz, nf2, nf1), dtype=self.cdtype, device=self.device)
        slab[:, self.idx[1][:, None], self.idx[2][None, :]] = slab_modes
        grid = self.grid
        grid[self.z0:self.z1] = self._fft(slab, (1, 2))
        if W > 1:
            # neighbours need my edge planes as their ghosts
            up = grid[self.z1 - self.below:self.z1].contiguous()    # next rank's planes below it
            down = grid[self.z0:self.z0 + self.above].contiguous()  # previous rank's planes above
            from_prev, from_next = self._ring(up, down)
            grid[self._planes(self.z0 - self.below, self.below)] = from_prev
            grid[self._planes(self.z1, self.above)] = from_next
        return self.local.execute(grid)

    def execute(self, data):
        """type 1: strengths of the local points -> fk[:, y_lo:y_hi, :] (ms3, my, ms1).
        type 2: that block of modes -> values at the local points."""
        return self._execute1(data) if self.type == 1 else self._execute2(data)

    def gather_modes(self, block):
        """Full (ms3, ms2, ms1) mode array on every rank from the per-rank blocks (tests)."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return block
        ymax = max(self.ysize)
        buf = torch.zeros((self.ms[0], ymax, self.ms[2]), dtype=self.cdtype, device=self.device)
        buf[:, :block.shape[1]] = block
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        return torch.cat([p[:, :n] for p, n in zip(parts, self.ysize)], dim=1)

    def destroy(self):
        if hasattr(self.local, "destroy"):
            self.local.destroy()


def route_points(z, y, x, c, nf3, group=None):
    """Exchange points (and strengths, or None) so that every rank ends with the points of its
    slab: one all_to_all of counts, then one all_to_all per array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return z, y, x, c
    owner = slab_of_points(z, nf3, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    ins, outs = counts.tolist(), recv_counts.tolist()
    res = []
    for a in (z, y, x, c):
        if a is None:
            res.append(None)
            continue
        src = a[order].contiguous()
        dst = torch.empty(int(sum(outs)), dtype=a.dtype, device=a.device)
        dist.all_to_all_single(dst, src, output_split_sizes=outs, input_split_sizes=ins,
                               group=group)
        res.append(dst)
    return tuple(res)
