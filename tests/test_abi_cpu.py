"""CPU-only: the C-ABI library loads, exports every symbol the headers in include/ declare,
its option structs have the reference layout, and argument validation returns the reference's
error codes before any CUDA work (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import finufft_b200
    return finufft_b200.load()


def _declared_symbols():
    """Every function name declared in include/*.h (after expanding the simple-call macros)."""
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", fn)).read()
        for pre_macro, stem in (("B200_CU_SIMPLE", "cufinufft"), ("B200_HOST_SIMPLE", "finufft")):
            if f"#define {pre_macro}(" in text:
                body = text.split(f"#define {pre_macro}(")[1].split(f"{pre_macro}(, double)")[0]
                for m in re.finditer(stem + r"##P##(\dd\d(?:many)?)\(", body):
                    for p in ("", "f"):
                        names.add(f"{stem}{p}{m.group(1)}")
        text = re.sub(r"#define B200_\w+_SIMPLE\(.*?\n(?=B200_)", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:int|void|int64_t|double|const char \*)\s*\*?(\w+)\(", text, re.M):
            names.add(m.group(1))
    return names


def test_exports_every_declared_symbol(lib):
    import finufft_b200
    declared = _declared_symbols()
    assert len(declared) >= 98
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and the python-side list is the same set
    assert set(finufft_b200.ALL_SYMBOLS) == declared


def test_opts_struct_layout_matches_reference():
    """Byte layout of cufinufft_opts (reference include/cufinufft_opts.h:7-40, mirrored by
    python/cufinufft/cufinufft/_cufinufft.py:90-110) and finufft_opts
    (include/finufft_opts.h:32-71) on LP64."""
    from finufft_b200._lib import CufinufftOpts, FinufftOpts
    assert C.sizeof(CufinufftOpts) == 88
    assert CufinufftOpts.upsampfac.offset == 0
    assert CufinufftOpts.gpu_method.offset == 8
    assert CufinufftOpts.gpu_maxsubprobsize.offset == 40
    assert CufinufftOpts.gpu_device_id.offset == 56
    assert CufinufftOpts.gpu_stream.offset == 64
    assert CufinufftOpts.modeord.offset == 72
    assert CufinufftOpts.debug.offset == 80
    assert C.sizeof(FinufftOpts) == 96
    assert FinufftOpts.upsampfac.offset == 40
    assert FinufftOpts.allow_eps_too_small.offset == 68
    assert FinufftOpts.fftw_lock_fun.offset == 72


def test_default_opts(lib):
    from finufft_b200._lib import CufinufftOpts, FinufftOpts
    o = CufinufftOpts()
    lib.cufinufft_default_opts(C.byref(o))
    # reference src/cuda/c_interface.cpp:153-172
    assert (o.upsampfac, o.gpu_method, o.gpu_sort, o.gpu_kerevalmeth, o.gpu_maxsubprobsize,
            o.gpu_device_id, o.modeord, o.gpu_spreadinterponly, o.debug) == \
        (0.0, 0, 1, 1, 1024, 0, 0, 0, 0)
    assert not o.gpu_stream
    h = FinufftOpts()
    lib.finufft_default_opts(C.byref(h))
    # reference include/finufft/plan.hpp:311-332
    assert (h.modeord, h.spreadinterponly, h.debug, h.showwarn, h.spread_sort, h.upsampfac,
            h.maxbatchsize, h.spread_nthr_atomic, h.allow_eps_too_small) == \
        (0, 0, 0, 1, 2, 0.0, 0, -1, 0)


def test_argument_validation_codes(lib):
    """reference src/cuda/c_interface.cpp:14-45,90-91,120-125 and src/cuda/makeplan.cu:256-264."""
    p = C.c_void_p()
    nm = (C.c_int64 * 3)(16, 16, 16)
    f = C.c_float(1e-4)
    assert lib.cufinufftf_makeplan(1, 4, nm, 1, 1, f, C.byref(p), None) == 12   # dim
    assert lib.cufinufftf_makeplan(1, 0, nm, 1, 1, f, C.byref(p), None) == 12
    assert lib.cufinufftf_makeplan(4, 3, nm, 1, 1, f, C.byref(p), None) == 10   # type
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 0, f, C.byref(p), None) == 9    # ntransf
    bad = (C.c_int64 * 3)(16, 0, 16)
    assert lib.cufinufftf_makeplan(1, 3, bad, 1, 1, f, C.byref(p), None) == 14  # n_modes <= 0
    big = (C.c_int64 * 3)(2 ** 31, 1, 1)
    assert lib.cufinufftf_makeplan(1, 1, big, 1, 1, f, C.byref(p), None) == 14  # > INT32_MAX
    prod = (C.c_int64 * 3)(2 ** 16, 2 ** 16, 1)
    assert lib.cufinufftf_makeplan(1, 2, prod, 1, 1, f, C.byref(p), None) == 14  # product
    assert lib.cufinufft_destroy(None) == 16     # GPU: ERR_PLAN_NOTVALID
    assert lib.cufinufftf_destroy(None) == 16
    assert lib.finufft_destroy(None) == 1        # CPU API: reference src/c_interface.cpp:94-95
    assert lib.b200_version().startswith(b"finufft_b200")


def test_no_cpu_fallback(lib):
    """Without a GPU a valid makeplan must fail loudly (code 15), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = C.c_void_p()
    nm = (C.c_int64 * 3)(16, 16, 16)
    assert lib.cufinufftf_makeplan(1, 3, nm, 1, 1, C.c_float(1e-4), C.byref(p), None) == 15
    assert not p.value
    hp = C.c_void_p()
    assert lib.finufft_makeplan(1, 3, nm, 1, 1, C.c_double(1e-6), C.byref(hp), None) == 15


def test_c_caller_compiles_links_and_gets_the_no_device_code(tmp_path):
    """A plain C99 program written against the reference's guru sequence (examples/guru3d1f.c)
    compiles with the headers in include/, links against the shared library, and -- on a machine
    without a CUDA device -- gets error 15 back from makeplan instead of any CPU fallback."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    import finufft_b200
    finufft_b200.load()
    libdir = os.path.join(ROOT, "finufft_b200")
    exe = str(tmp_path / "guru3d1f")
    subprocess.check_call(["gcc", "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "guru3d1f.c"), "-L", libdir,
                           "-lfinufft_b200", "-lm", f"-Wl,-rpath,{libdir}", "-o", exe])
    # headers are valid C++ too
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text('#include "b200_finufft.h"\n#include "b200_cufinufft.h"\n'
                   '#include "b200_introspect.h"\n#include "b200_sharded.h"\n'
                   'int main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           str(cpp)])
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    rc = subprocess.call([exe])
    assert rc == (0 if has_gpu else 15)
    # the sharded entry points from C (examples/sharded3d1f.c), when the CUDA runtime is there
    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        exe2 = str(tmp_path / "sharded3d1f")
        subprocess.check_call(["gcc", "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-Werror", "-I",
                               os.path.join(ROOT, "include"), "-I", cuda_inc,
                               os.path.join(ROOT, "examples", "sharded3d1f.c"), "-L", libdir,
                               "-lfinufft_b200", "-L", cuda_lib, "-lcudart", "-lm",
                               f"-Wl,-rpath,{libdir}", "-o", exe2])
        rc = subprocess.call([exe2])
        assert rc == (0 if has_gpu else 15)
