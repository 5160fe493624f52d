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
    if force or _stale(out, [src, os.path.join(HERE, "fft_standin.hpp")]):
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


def build_ref_library(force=False, verbose=False):
    """The reference's own CPU library (finufft[f]_makeplan / setpts / execute / destroy and the
    simple interfaces): src/*.cpp, src/common/*.cpp and every include/finufft/*.hpp compiled
    from where they lie under /root/reference, twice for the precision-specific sources (with
    and without FINUFFT_SINGLE, src/CMakeLists.txt:17-34,80-82), with the reference's own
    release flags (cmake/toolchain.cmake:12-24) and -mavx2 -mfma in place of -march=native so
    the library also runs on the GPU box's host CPU.  Its three un-vendored third-party
    dependencies are replaced by the stand-ins under oracle/shim/: xsimd (SIMD wrapper, lane
    order only), POET (compile-time dispatch, no arithmetic) and FFTW (the FFT itself is then
    oracle/fft_standin.hpp).  Sort, spread, interp, deconvolve, type-3 set-up and the guru
    driver are the reference's code, unmodified.  -> oracle/_ref/libfinufft_ref.so or None."""
    outdir = os.path.join(HERE, "_ref")
    objdir = os.path.join(outdir, "obj")
    out = os.path.join(outdir, "libfinufft_ref.so")
    src = os.path.join(REF, "src")
    shim = os.path.join(HERE, "shim")
    per_prec = ["makeplan.cpp", "setpts.cpp", "execute.cpp", "spreadinterp.cpp",
                "spreadinterp_1d.cpp", "spreadinterp_2d.cpp", "spreadinterp_3d.cpp"]
    common = ["fft.cpp", "c_interface.cpp", "utils.cpp", "common/kernel.cpp", "common/pswf.cpp",
              "common/utils.cpp"]
    if not all(os.path.exists(os.path.join(src, f)) for f in per_prec + common):
        return out if os.path.exists(out) else None
    shims = [os.path.join(shim, f) for f in ("xsimd/xsimd.hpp", "poet/poet.hpp", "fftw3.h",
                                             "fftw_standin.cpp")] + [
        os.path.join(HERE, "fft_standin.hpp"), os.path.join(HERE, "ref_lib_shim.cpp")]
    inc = os.path.join(REF, "include")
    hdrs = [os.path.join(inc, "finufft", f) for f in os.listdir(os.path.join(inc, "finufft"))]
    if not (force or _stale(out, shims + hdrs + [os.path.join(src, f) for f in per_prec + common])):
        return out
    os.makedirs(objdir, exist_ok=True)
    flags = ["-std=c++17", "-O3", "-funroll-loops", "-ffp-contract=fast", "-fno-math-errno",
             "-fno-signed-zeros", "-fno-trapping-math", "-fassociative-math", "-freciprocal-math",
             "-fmerge-all-constants", "-ftree-vectorize", "-fimplicit-constexpr",
             "-fcx-limited-range", "-fno-semantic-interposition", "-mavx2", "-mfma", "-fopenmp",
             "-fPIC", "-fvisibility=hidden", "-DFINUFFT_NO_DEPRECATED_FIELDS", "-DFINUFFT_DLL",
             "-Ddll_EXPORTS", "-I", shim, "-I", inc]
    jobs = []
    for f in per_prec:
        jobs.append((os.path.join(src, f), os.path.join(objdir, f[:-4] + "_f64.o"), []))
        jobs.append((os.path.join(src, f), os.path.join(objdir, f[:-4] + "_f32.o"),
                     ["-DFINUFFT_SINGLE"]))
    for f in common:
        jobs.append((os.path.join(src, f),
                     os.path.join(objdir, f.replace("/", "_")[:-4] + ".o"), []))
    jobs.append((os.path.join(shim, "fftw_standin.cpp"), os.path.join(objdir, "fftw_standin.o"),
                 []))
    jobs.append((os.path.join(HERE, "ref_lib_shim.cpp"), os.path.join(objdir, "ref_lib_shim.o"),
                 []))

    def one(job):
        s, o, extra = job
        cmd = ["g++"] + flags + extra + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return o

    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, jobs))
    cmd = ["g++", "-shared", "-fopenmp", "-o", out] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


REF_API_TESTS = {
    # program: (source under /root/reference/test/cuda, compiler)
    "test_makeplan": ("test_makeplan.c", "gcc"),
    "public_api_test": ("public_api_test.c", "gcc"),
    "cufinufft_error_handling": ("cufinufft_error_handling.cu", "nvcc"),
    "cufinufft_multigpu_test": ("cufinufft_multigpu_test.cu", "nvcc"),
    "cufinufft_simple_test": ("cufinufft_simple_test.cu", "nvcc"),
}


def build_ref_api_tests(force=False, verbose=False):
    """The reference's own C / CUDA API contract tests (test/cuda/*.c, *.cu: plain programs over
    <cufinufft.h>), compiled from the sources where they lie and LINKED AGAINST OUR LIBRARY
    (finufft_b200/libfinufft_b200.so in place of libcufinufft), into oracle/_ref/bin/.  They run
    on the GPU box from tests/test_gpu_ref_api.py.  The accuracy programs cufinufft{1,2,3}d_test.cu
    include the reference's internal headers, which need the un-vendored POET dependency, and
    cannot be built here.  Returns the list of programs present."""
    outdir = os.path.join(HERE, "_ref", "bin")
    srcdir = os.path.join(REF, "test", "cuda")
    lib = os.path.join(HERE, "..", "finufft_b200", "libfinufft_b200.so")
    have = []
    for prog, (src, cc) in REF_API_TESTS.items():
        out = os.path.join(outdir, prog)
        path = os.path.join(srcdir, src)
        if not os.path.exists(path) or not os.path.exists(lib):
            if os.path.exists(out):
                have.append(out)
            continue
        if force or _stale(out, [path]):
            os.makedirs(outdir, exist_ok=True)
            rpath = "$ORIGIN/../../../finufft_b200"
            if cc == "gcc":
                cmd = ["gcc", "-O1", "-I", os.path.join(REF, "include"), "-I",
                       "/usr/local/cuda/include", path, "-o", out, os.path.abspath(lib),
                       "-L/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{rpath}"]
            else:
                cmd = ["/usr/local/cuda/bin/nvcc", "-O1", "-std=c++17", "-gencode",
                       "arch=compute_100a,code=sm_100a", "-I", os.path.join(REF, "include"), path,
                       "-o", out, os.path.abspath(lib), "-Xlinker", f"-rpath={rpath}"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        have.append(out)
    return have


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv, verbose=True))
    print(build_ref(force="--force" in sys.argv, verbose=True))
    print(build_ref_dirft(force="--force" in sys.argv, verbose=True))
    print(build_ref_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_ref_api_tests(force="--force" in sys.argv, verbose=True))
