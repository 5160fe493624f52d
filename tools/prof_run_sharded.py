"""Profiler driver for the sharded plan at world = 1 (no exchanges): C3-size transform through
b200_slabf_* so that ncu lists the slab kernels (pack, deconvolve, 2D / 1D cuFFT) next to the
spread kernel.  python tools/prof_run_sharded.py [M]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import perfdata
from finufft_b200.sharded import ShardedPlan

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
modes = (256, 256, 256)
for type_ in (1, 2):
    sp = ShardedPlan(type_, modes, eps=1e-6, dtype="complex64", upsampfac=2.0)
    x, y, z = (torch.from_numpy(p).cuda() for p in perfdata.points(3, M, np.float32))
    sp.setpts(z, y, x)
    n_in = M if type_ == 1 else 256 ** 3
    data = torch.from_numpy(perfdata.strengths(n_in, np.complex64).reshape((M,) if type_ == 1 else sp.block_shape)).cuda()
    out = sp.execute(data)
    out = sp.execute(data, out)
    torch.cuda.synchronize()
    print("type", type_, {k: round(v, 3) for k, v in sp.stage_ms().items()})
    sp.destroy()
