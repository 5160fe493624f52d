import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (test infrastructure only)."""
    from oracle import build as obuild
    obuild.build_oracle()
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import finufft_b200
    finufft_b200.load()  # must exist: no fallback
    return torch


def make_points(rng, dim, M, rt, kind="uniform", nf=None):
    """uniform: iid in [-pi,pi); cluster: all points inside an 8-cell cube of the fine grid
    (SURVEY.md 8d definition); wide: far outside [-pi,pi) to exercise folding."""
    out = []
    for d in range(dim):
        if kind == "uniform":
            a = rng.uniform(-np.pi, np.pi, M)
        elif kind == "cluster":
            h = 2 * np.pi / nf[d]
            a = rng.uniform(0, 8 * h, M)
        elif kind == "wide":
            a = rng.uniform(-40.0, 40.0, M)
        elif kind == "edges":
            a = rng.choice(np.array([-np.pi, np.pi, 0.0, -3.1415925, 3.1415925, 1e-7, -1e-7]), M)
        else:
            raise ValueError(kind)
        out.append(a.astype(rt))
    return out + [None] * (3 - dim)
