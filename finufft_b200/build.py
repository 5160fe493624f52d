"""Build finufft_b200/libfinufft_b200.so for sm_100a with plain nvcc (no CMake, no JIT cache).

    python -m finufft_b200.build [--force] [-v]

Objects go to finufft_b200/build/, the shared library stays in-tree next to this file so it
travels with the repository snapshot to the GPU box.  Only files whose sources changed are
recompiled; translation units are compiled in parallel.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfinufft_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler",
          "-fvisibility=hidden"]
# capi.cu carries the exported symbols
EXPORT = ["-Xcompiler", "-fvisibility=default"]

SOURCES = [
    "planmath.cpp", "sort.cu", "partition.cu", "gridops.cu", "engine.cu", "sweep3d.cu", "stage.cu", "sweep2d_f32.cu", "sweep2d_f64.cu", "type3.cu", "direct.cu", "slab.cu", "capi.cu", "capi_sharded.cu",
    "spreadinterp_f32_d1.cu", "spreadinterp_f32_d2.cu", "spreadinterp_f32_d3.cu",
    "spreadinterp_f64_d1.cu", "spreadinterp_f64_d2.cu", "spreadinterp_f64_d3.cu",
]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    inc = os.path.join(HERE, "..", "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    return hs


def _needs(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    cmd = [NVCC] + ARCH + COMMON
    if src in ("capi.cu", "capi_sharded.cu"):
        cmd += EXPORT
    if src.endswith(".cpp"):
        cmd += ["-Xcompiler", "-ffp-contract=off", "-x", "cu"]
    cmd += ["-c", path, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo, objs = [], []
    for s in SOURCES:
        obj = os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _needs(obj, [os.path.join(CSRC, s)] + hdrs):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcufft", "-ldl", "-Xlinker",
                                                               "--no-undefined", "-Xlinker",
                                                               "-soname=libfinufft_b200.so"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
