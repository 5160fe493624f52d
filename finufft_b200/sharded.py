"""Host-side binding of the sharded (multi-GPU) plan, include/b200_sharded.h.

``ShardedPlan`` is one large 3D type-1 / type-2 transform spread over the GPUs of one box by
z-slabs of the fine grid; all of it (routing of the points, spreading into the slab window,
ghost-plane exchange, 2D FFT, slab<->pencil transpose, 1D FFT, deconvolution) runs inside
libfinufft_b200.so with NCCL.  `torch.distributed` is used here for one thing only: handing
the 128-byte NCCL unique id from rank 0 to the other ranks.

Array conventions follow ``finufft_b200.Plan`` (python/cufinufft/cufinufft/_plan.py:211-237):
C order, so n_modes = (ms3, ms2, ms1) and setpts takes (z, y, x); the mode block of a rank is
fk[:, ylo:yhi, :].
"""
import ctypes as C

import numpy as np

from . import _lib
from .plan import _check, _dtypes


def _share_unique_id(lib, rank, world, group):
    """rank 0 creates the id; torch.distributed (any backend) carries it to the others."""
    buf = (C.c_ubyte * 128)()
    if world == 1:
        return buf
    import torch
    import torch.distributed as dist
    if rank == 0:
        _check(lib.b200_slab_unique_id(C.byref(buf)), "unique_id")
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else "cpu"
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().tolist())
    return (C.c_ubyte * 128).from_buffer_copy(raw)


class ShardedPlan:
    """3D transform of type 1 or 2 sharded over the ranks of a torch.distributed group.

    Every rank constructs the plan (collective), passes its own points to setpts and its own
    strengths / mode block to execute.  Keyword options are fields of cufinufft_opts
    (upsampfac, modeord, gpu_device_id, gpu_stream, ...).
    """

    def __init__(self, nufft_type, n_modes, eps=1e-6, isign=None, dtype="complex64", group=None,
                 **kwargs):
        import torch
        import torch.distributed as dist
        self._torch = torch
        self._lib = _lib.load()
        self._pre, self._real, self._cplx = _dtypes(dtype)
        if nufft_type not in (1, 2) or len(n_modes) != 3:
            raise ValueError("ShardedPlan handles 3D transforms of type 1 and 2")
        self.type = nufft_type
        self.n_modes = tuple(int(n) for n in n_modes)
        self.isign = (1 if nufft_type == 1 else -1) if isign is None else isign
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        opts = _lib.CufinufftOpts()
        self._lib.cufinufft_default_opts(C.byref(opts))
        kwargs.setdefault("gpu_device_id", torch.cuda.current_device())
        for k, v in kwargs.items():
            if not hasattr(opts, k):
                raise TypeError(f"invalid option '{k}'")
            setattr(opts, k, v)
        self.device = torch.device("cuda", opts.gpu_device_id)
        uid = _share_unique_id(self._lib, self.rank, self.world, group)
        nm = (C.c_int64 * 3)(*self.n_modes[::-1])
        self._plan = C.c_void_p()
        real = C.c_float if self._pre == "f" else C.c_double
        mk = getattr(self._lib, f"b200_slab{self._pre}_makeplan")
        _check(mk(nufft_type, nm, self.isign, real(eps), self.rank, self.world, C.byref(uid),
                  C.byref(opts), C.byref(self._plan)), "slab_makeplan")
        i = self.info()
        self.y_lo, self.y_hi = i["ylo"], i["yhi"]
        self.block_shape = (self.n_modes[0], self.y_hi - self.y_lo, self.n_modes[2])
        self.M = 0
        self._refs = []

    def info(self):
        out = _lib.SlabInfo()
        _check(self._lib.b200_slab_get_info(self._plan, C.byref(out)), "slab_info")
        return dict(ns=out.ns, mode=out.mode, nf=list(out.nf), ms=list(out.ms), z0=out.z0,
                    nz=out.nz, ylo=out.ylo, yhi=out.yhi, win_org=out.win_org, win_n=out.win_n,
                    M=out.M, M_local=out.M_local, rank=out.rank, world=out.world)

    def _real_tensor(self, a, name):
        torch = self._torch
        want = torch.float32 if self._pre == "f" else torch.float64
        if not (torch.is_tensor(a) and a.is_cuda and a.dtype == want):
            raise TypeError(f"{name} must be a CUDA tensor of dtype {want}")
        return a.contiguous()

    def setpts(self, z, y, x, routed=False):
        """This rank's points (python order: slowest axis first).  routed=True promises that
        every z folds into this rank's slab of the fine grid."""
        z, y, x = (self._real_tensor(a, n) for a, n in zip((z, y, x), "zyx"))
        M = x.numel()
        if y.numel() != M or z.numel() != M:
            raise TypeError("coordinate arrays must have equal length")
        sp = getattr(self._lib, f"b200_slab{self._pre}_setpts")
        _check(sp(self._plan, M, x.data_ptr(), y.data_ptr(), z.data_ptr(), int(bool(routed))),
               "slab_setpts")
        self._refs = [x, y, z]
        self.M = M

    def execute(self, data, out=None):
        """type 1: strengths (M,) -> mode block (ms3, yhi-ylo, ms1); type 2: the reverse."""
        torch = self._torch
        want = torch.complex64 if self._pre == "f" else torch.complex128
        if not (torch.is_tensor(data) and data.is_cuda and data.dtype == want):
            raise TypeError(f"data must be a CUDA tensor of dtype {want}")
        data = data.contiguous()
        in_shape, out_shape = ((self.M,), self.block_shape) if self.type == 1 else (
            self.block_shape, (self.M,))
        if tuple(data.shape) != in_shape:
            raise TypeError(f"data must have shape {in_shape}, got {tuple(data.shape)}")
        if out is None:
            out = torch.empty(out_shape, dtype=want, device=data.device)
        ex = getattr(self._lib, f"b200_slab{self._pre}_execute")
        c, fk = (data, out) if self.type == 1 else (out, data)
        _check(ex(self._plan, c.data_ptr(), fk.data_ptr()), "slab_execute")
        return out

    def gather_modes(self, block):
        torch = self._torch
        full = torch.empty(self.n_modes, dtype=block.dtype, device=block.device)
        f = getattr(self._lib, f"b200_slab{self._pre}_gather_modes")
        _check(f(self._plan, block.contiguous().data_ptr(), full.data_ptr()), "slab_gather")
        return full

    def slice_modes(self, full):
        torch = self._torch
        block = torch.empty(self.block_shape, dtype=full.dtype, device=full.device)
        f = getattr(self._lib, f"b200_slab{self._pre}_slice_modes")
        _check(f(self._plan, full.contiguous().data_ptr(), block.data_ptr()), "slab_slice")
        return block

    def stage_ms(self):
        ms = (C.c_float * 10)()
        _check(self._lib.b200_slab_get_stage_ms(self._plan, C.byref(ms)), "slab_stage_ms")
        keys = ("spreadinterp", "ghost", "fft2d", "pack", "transpose", "fft1d", "deconv",
                "route_values", "execute", "setpts")
        return {k: float(ms[i]) for i, k in enumerate(keys)}

    def launch_count(self):
        n = C.c_uint64()
        _check(self._lib.b200_slab_get_launch_count(self._plan, C.byref(n)), "slab_launches")
        return int(n.value)

    def destroy(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            getattr(self._lib, f"b200_slab{self._pre}_destroy")(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
