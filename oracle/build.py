"""ORACLE — TEST INFRASTRUCTURE ONLY.  Build recipes for the CPU checker.

  python oracle/build.py            # builds oracle/liboracle.so (+ oracle/_ref if possible)

* oracle/liboracle.so   — our CPU restatement (oracle/finufft_oracle.cpp), strict IEEE:
                          -O2 -fno-fast-math -ffp-contract=off, OpenMP.
* oracle/_ref/libfinufft_ref_common.so — the REFERENCE's own src/common/{kernel,pswf,utils}.cpp
                          compiled where they lie under /root/reference with oracle/ref_shim.cpp
                          on top.  Only built when /root/reference exists (this container);
                          the GPU box uses the prebuilt file that travels with the snapshot.
* oracle/_ref/libfinufft_ref_dirft.so — the REFERENCE's own direct sums and error norms
                          (test/utils/dirft{1,2,3}d.hpp, norms.hpp) behind oracle/ref_dirft_shim.cpp:
                          the direct-sum checker of every accuracy test.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False, verbose=False):
    src = os.path.join(HERE, "finufft_oracle.cpp")
    out = os.path.join(HERE, "liboracle.so")
    if force or _stale(out, [src]):
        cmd = ["g++", "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fopenmp",
               "-shared", "-fPIC", src, "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def build_ref(force=False, verbose=False):
    """Returns the path of the reference-common library, or None if it cannot be had."""
    outdir = os.path.join(HERE, "_ref")
    out = os.path.join(outdir, "libfinufft_ref_common.so")
    shim = os.path.join(HERE, "ref_shim.cpp")
    srcs = [os.path.join(REF, "src", "common", f) for f in ("kernel.cpp", "pswf.cpp", "utils.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return out if os.path.exists(out) else None
    if force or _stale(out, [shim] + srcs):
        os.makedirs(outdir, exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", os.path.join(REF, "include"),
               shim] + srcs + ["-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def build_ref_dirft(force=False, verbose=False):
    """The reference's own direct-sum checkers and norms (test/utils/dirft{1,2,3}d.hpp,
    norms.hpp: dependency-free templates) compiled where they lie, with oracle/ref_dirft_shim.cpp
    on top, into oracle/_ref/libfinufft_ref_dirft.so.  Returns the path or None."""
    outdir = os.path.join(HERE, "_ref")
    out = os.path.join(outdir, "libfinufft_ref_dirft.so")
    shim = os.path.join(HERE, "ref_dirft_shim.cpp")
    inc = os.path.join(REF, "test", "utils")
    hdrs = [os.path.join(inc, f) for f in ("dirft1d.hpp", "dirft2d.hpp", "dirft3d.hpp",
                                            "norms.hpp")]
    if not all(os.path.exists(h) for h in hdrs):
        return out if os.path.exists(out) else None
    if force or _stale(out, [shim] + hdrs):
        os.makedirs(outdir, exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-fopenmp", "-shared", "-fPIC", "-I", inc, shim,
               "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv, verbose=True))
    print(build_ref(force="--force" in sys.argv, verbose=True))
    print(build_ref_dirft(force="--force" in sys.argv, verbose=True))
