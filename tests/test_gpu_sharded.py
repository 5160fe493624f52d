"""The sharded (z-slab) plan, include/b200_sharded.h, on real GPUs: world 1 in this process's
box whatever it has, world 2 when two GPUs are visible (tools/sharded_check.py under torchrun,
NCCL).  The checker compares with the unsharded plan: uniform, clustered (replicated-window
mode) and pre-partitioned points, both types, both precisions, FFT mode order, odd sizes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(cmd):
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(p.stdout[-4000:])
    print(p.stderr[-3000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "sharded_check:" in p.stdout


def test_sharded_world1(cuda):
    _run([sys.executable, "tools/sharded_check.py"])


def test_sharded_world2_nccl(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
          "--master-addr", "127.0.0.1", "--master-port", "29617", "tools/sharded_check.py"])
