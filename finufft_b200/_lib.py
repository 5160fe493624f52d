"""ctypes binding of finufft_b200/libfinufft_b200.so (the C ABI in include/b200_*.h).

This is the binding a maintainer of the reference's python packages would write against the
drop-in library: the structures mirror python/cufinufft/cufinufft/_cufinufft.py:90-110
(cufinufft_opts) and python/finufft/finufft/_finufft.py (finufft_opts) field for field.
There is NO fallback: if the CUDA library is missing, importing fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200_NUFFT_LIBRARY: development hook (kernel-variant A/B probes); unset in normal use
LIB_PATH = os.environ.get("B200_NUFFT_LIBRARY") or os.path.join(HERE, "libfinufft_b200.so")


class CufinufftOpts(C.Structure):
    _fields_ = [
        ("upsampfac", C.c_double),
        ("gpu_method", C.c_int),
        ("gpu_sort", C.c_int),
        ("gpu_binsizex", C.c_int),
        ("gpu_binsizey", C.c_int),
        ("gpu_binsizez", C.c_int),
        ("gpu_obinsizex", C.c_int),
        ("gpu_obinsizey", C.c_int),
        ("gpu_obinsizez", C.c_int),
        ("gpu_maxsubprobsize", C.c_int),
        ("gpu_kerevalmeth", C.c_int),
        ("gpu_spreadinterponly", C.c_int),
        ("gpu_maxbatchsize", C.c_int),
        ("gpu_device_id", C.c_int),
        ("gpu_stream", C.c_void_p),
        ("modeord", C.c_int),
        ("gpu_np", C.c_int),
        ("debug", C.c_int),
    ]


class FinufftOpts(C.Structure):
    _fields_ = [
        ("modeord", C.c_int),
        ("spreadinterponly", C.c_int),
        ("debug", C.c_int),
        ("spread_debug", C.c_int),
        ("showwarn", C.c_int),
        ("nthreads", C.c_int),
        ("fftw", C.c_int),
        ("spread_sort", C.c_int),
        ("spread_kerevalmeth", C.c_int),
        ("spread_kerpad", C.c_int),
        ("upsampfac", C.c_double),
        ("spread_thread", C.c_int),
        ("maxbatchsize", C.c_int),
        ("spread_nthr_atomic", C.c_int),
        ("spread_max_sp_size", C.c_int),
        ("spread_kerformula", C.c_int),
        ("allow_eps_too_small", C.c_int),
        ("fftw_lock_fun", C.c_void_p),
        ("fftw_unlock_fun", C.c_void_p),
        ("fftw_lock_data", C.c_void_p),
    ]


class SlabInfo(C.Structure):
    _fields_ = [
        ("is_float", C.c_int), ("type", C.c_int), ("rank", C.c_int), ("world", C.c_int),
        ("ns", C.c_int), ("mode", C.c_int),
        ("nf", C.c_int64 * 3), ("ms", C.c_int64 * 3),
        ("z0", C.c_int64), ("nz", C.c_int64), ("ylo", C.c_int64), ("yhi", C.c_int64),
        ("win_org", C.c_int64), ("win_n", C.c_int64), ("M", C.c_int64), ("M_local", C.c_int64),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("is_float", C.c_int), ("type", C.c_int), ("dim", C.c_int), ("ntr", C.c_int),
        ("ns", C.c_int), ("nc", C.c_int), ("batch", C.c_int),
        ("sigma", C.c_double), ("beta", C.c_double), ("tol", C.c_double),
        ("nf", C.c_int64 * 3), ("ms", C.c_int64 * 3), ("nbins", C.c_int64 * 3),
        ("M", C.c_int64), ("nsub", C.c_int64),
    ]


# every symbol the headers in include/ declare (checked by tests/test_abi.py)
GURU_GPU = ["cufinufft_default_opts"] + [
    f"cufinufft{p}_{n}" for p in ("", "f") for n in ("makeplan", "setpts", "execute", "destroy")]
GURU_HOST = [f"finufft{p}_{n}" for p in ("", "f")
             for n in ("default_opts", "makeplan", "setpts", "execute", "execute_adjoint",
                       "destroy")]
SIMPLE_GPU = [f"cufinufft{p}{d}d{t}{m}" for p in ("", "f") for d in (1, 2, 3)
              for t in (1, 2, 3) for m in ("", "many")]
SIMPLE_HOST = [f"finufft{p}{d}d{t}{m}" for p in ("", "f") for d in (1, 2, 3)
               for t in (1, 2, 3) for m in ("", "many")]
INTROSPECT = ["b200_get_plan_info", "b200_get_inner_plan_info", "b200_host_sigma_candidates",
              "b200_host_choose_sigma_type3", "b200_get_sort_permutation", "b200_get_raw_sort_order",
              "b200_get_sort_path", "b200_get_window_table",
              "b200_get_phihat", "b200_enable_profiling", "b200_get_stage_ms",
              "b200_get_launch_count", "b200_host_kernel", "b200_host_fine_grid",
              "b200_host_fseries", "b200_host_smallest_sigma", "b200_host_sigma_feasible",
              "b200_host_choose_sigma", "b200_version"]
SHARDED = ["b200_slab_unique_id", "b200_slab_get_info", "b200_slab_get_stage_ms",
           "b200_slab_get_launch_count"] + [
    f"b200_slab{p}_{n}" for p in ("", "f")
    for n in ("makeplan", "setpts", "execute", "gather_modes", "slice_modes", "destroy")]
ALL_SYMBOLS = GURU_GPU + GURU_HOST + SIMPLE_GPU + SIMPLE_HOST + INTROSPECT + SHARDED

_lib = None


def load():
    """Load the CUDA library. Raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m finufft_b200.build` "
            "(finufft_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    i64, vp, dbl, flt, ci = C.c_int64, C.c_void_p, C.c_double, C.c_float, C.c_int
    lib.cufinufft_default_opts.argtypes = [C.POINTER(CufinufftOpts)]
    lib.cufinufft_default_opts.restype = None
    for pre, real in (("", dbl), ("f", flt)):
        mk = getattr(lib, f"cufinufft{pre}_makeplan")
        mk.argtypes = [ci, ci, C.POINTER(i64), ci, ci, real, C.POINTER(vp), C.POINTER(CufinufftOpts)]
        mk.restype = ci
        sp = getattr(lib, f"cufinufft{pre}_setpts")
        sp.argtypes = [vp, i64, vp, vp, vp, ci, vp, vp, vp]
        sp.restype = ci
        ex = getattr(lib, f"cufinufft{pre}_execute")
        ex.argtypes = [vp, vp, vp]
        ex.restype = ci
        de = getattr(lib, f"cufinufft{pre}_destroy")
        de.argtypes = [vp]
        de.restype = ci
        do = getattr(lib, f"finufft{pre}_default_opts")
        do.argtypes = [C.POINTER(FinufftOpts)]
        do.restype = None
        mk = getattr(lib, f"finufft{pre}_makeplan")
        mk.argtypes = [ci, ci, C.POINTER(i64), ci, ci, real, C.POINTER(vp), C.POINTER(FinufftOpts)]
        mk.restype = ci
        sp = getattr(lib, f"finufft{pre}_setpts")
        sp.argtypes = [vp, i64, vp, vp, vp, i64, vp, vp, vp]
        sp.restype = ci
        for nm in ("execute", "execute_adjoint"):
            ex = getattr(lib, f"finufft{pre}_{nm}")
            ex.argtypes = [vp, vp, vp]
            ex.restype = ci
        de = getattr(lib, f"finufft{pre}_destroy")
        de.argtypes = [vp]
        de.restype = ci
    lib.b200_get_plan_info.argtypes = [vp, C.POINTER(PlanInfo)]
    lib.b200_get_plan_info.restype = ci
    lib.b200_get_inner_plan_info.argtypes = [vp, C.POINTER(PlanInfo)]
    lib.b200_get_inner_plan_info.restype = ci
    lib.b200_get_sort_permutation.argtypes = [vp, vp]
    lib.b200_get_sort_permutation.restype = ci
    lib.b200_get_raw_sort_order.argtypes = [vp, vp]
    lib.b200_get_raw_sort_order.restype = ci
    lib.b200_get_sort_path.argtypes = [vp, C.POINTER(ci)]
    lib.b200_get_sort_path.restype = ci
    lib.b200_get_window_table.argtypes = [vp, vp]
    lib.b200_get_window_table.restype = ci
    lib.b200_get_phihat.argtypes = [vp, ci, vp]
    lib.b200_get_phihat.restype = ci
    lib.b200_enable_profiling.argtypes = [vp, ci]
    lib.b200_enable_profiling.restype = ci
    lib.b200_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 5)]
    lib.b200_get_stage_ms.restype = ci
    lib.b200_get_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.b200_get_launch_count.restype = ci
    lib.b200_host_kernel.argtypes = [dbl, ci, ci, dbl, ci, ci, C.POINTER(ci), C.POINTER(dbl),
                                     C.POINTER(ci), vp]
    lib.b200_host_kernel.restype = ci
    lib.b200_host_fine_grid.argtypes = [dbl, i64, ci]
    lib.b200_host_fine_grid.restype = i64
    lib.b200_host_smallest_sigma.argtypes = [dbl, ci, ci, ci, dbl]
    lib.b200_host_smallest_sigma.restype = dbl
    lib.b200_host_sigma_feasible.argtypes = [dbl, dbl, ci, ci, ci, dbl]
    lib.b200_host_sigma_feasible.restype = ci
    lib.b200_host_choose_sigma.argtypes = [dbl, ci, ci, ci, C.POINTER(i64), dbl]
    lib.b200_host_choose_sigma.restype = dbl
    lib.b200_host_sigma_candidates.argtypes = [dbl, ci, ci, ci, dbl, dbl, C.POINTER(dbl),
                                               C.POINTER(ci), ci]
    lib.b200_host_sigma_candidates.restype = ci
    lib.b200_host_choose_sigma_type3.argtypes = [dbl, ci, ci, dbl, dbl, C.POINTER(dbl),
                                                 C.POINTER(dbl)]
    lib.b200_host_choose_sigma_type3.restype = dbl
    lib.b200_host_fseries.argtypes = [i64, ci, ci, ci, vp, vp]
    lib.b200_host_fseries.restype = ci
    lib.b200_version.restype = C.c_char_p
    lib.b200_slab_unique_id.argtypes = [vp]
    lib.b200_slab_unique_id.restype = ci
    for pre, real in (("", dbl), ("f", flt)):
        mk = getattr(lib, f"b200_slab{pre}_makeplan")
        mk.argtypes = [ci, C.POINTER(i64), ci, real, ci, ci, vp, C.POINTER(CufinufftOpts),
                       C.POINTER(vp)]
        mk.restype = ci
        sp = getattr(lib, f"b200_slab{pre}_setpts")
        sp.argtypes = [vp, i64, vp, vp, vp, ci]
        sp.restype = ci
        for nm in ("execute", "gather_modes", "slice_modes"):
            f = getattr(lib, f"b200_slab{pre}_{nm}")
            f.argtypes = [vp, vp, vp]
            f.restype = ci
        de = getattr(lib, f"b200_slab{pre}_destroy")
        de.argtypes = [vp]
        de.restype = ci
    lib.b200_slab_get_info.argtypes = [vp, C.POINTER(SlabInfo)]
    lib.b200_slab_get_info.restype = ci
    lib.b200_slab_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 10)]
    lib.b200_slab_get_stage_ms.restype = ci
    lib.b200_slab_get_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.b200_slab_get_launch_count.restype = ci
    _lib = lib
    return lib
