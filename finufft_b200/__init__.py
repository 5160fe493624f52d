"""finufft_b200 — a B200-native (sm_100a) NUFFT engine behind FINUFFT's guru-plan C ABI.

The product is finufft_b200/libfinufft_b200.so (hand-written CUDA kernels + cuFFT); this
package is the thin host-side mirror of the reference's Python interfaces.  There is no CPU
fallback: without the built library and a GPU, plans cannot be made.
"""
from ._lib import ALL_SYMBOLS, LIB_PATH, load  # noqa: F401
from .plan import HostPlan, NufftError, Plan  # noqa: F401
from .simple import (nufft1d1, nufft1d2, nufft1d3, nufft2d1, nufft2d2, nufft2d3,  # noqa: F401
                     nufft3d1, nufft3d2, nufft3d3)

__version__ = "0.1.0"
