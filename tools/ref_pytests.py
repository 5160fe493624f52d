"""Run the reference's OWN Python test suites against libfinufft_b200.so.

The reference's `python/cufinufft` and `python/finufft` packages are thin ctypes bindings that
look for `libcufinufft.so` / `libfinufft.so` next to their `__init__.py`
(python/cufinufft/cufinufft/_cufinufft.py:39-87, python/finufft/finufft/_finufft.py:48-90).  Our
library exports exactly those symbol names, so dropping it in under those file names makes the
reference's bindings -- and therefore its pytest suites -- run on the B200 engine unmodified.

    python tools/ref_pytests.py stage      # here (needs /root/reference): copies the two packages
                                           # and their tests into oracle/_ref/py/ (git-ignored,
                                           # travels to the GPU box like oracle/_ref's .so files)
    python tools/ref_pytests.py run [-k expr] [--which cufinufft|finufft|both]
                                           # on the GPU box: link our .so in, run pytest, write
                                           # gpurun_out/ref_pytests_<pkg>.log

Nothing from the reference is committed: oracle/_ref/ is in .gitignore.
"""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref", "py")
REF = "/root/reference/python"
LIB = os.path.join(ROOT, "finufft_b200", "libfinufft_b200.so")

PKGS = {
    # name: (package dir, tests dir, library file name inside the package, extra pytest args)
    "cufinufft": ("cufinufft/cufinufft", "cufinufft/tests", "libcufinufft.so",
                  ["--framework", "torch"]),
    "finufft": ("finufft/finufft", "finufft/test", "libfinufft.so", []),
}


def stage():
    if not os.path.isdir(REF):
        raise SystemExit("stage: /root/reference is not present (run this in the build container)")
    for name, (pkg, tests, _, _) in PKGS.items():
        dst = os.path.join(STAGE, name)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(REF, pkg), os.path.join(dst, name))
        shutil.copytree(os.path.join(REF, tests), os.path.join(dst, "tests"))
    ex = os.path.join(REF, "cufinufft", "examples")
    if os.path.isdir(ex):
        shutil.copytree(ex, os.path.join(STAGE, "cufinufft", "examples"), dirs_exist_ok=True)
    print("staged into", STAGE)


def run(which, kexpr, extra):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rc_all = 0
    for name in (PKGS if which == "both" else [which]):
        pkg, tests, libname, args = PKGS[name]
        base = os.path.join(STAGE, name)
        if not os.path.isdir(base):
            raise SystemExit(f"run: {base} missing; run `stage` first")
        shutil.copyfile(LIB, os.path.join(base, name, libname))
        env = dict(os.environ)
        env["PYTHONPATH"] = base + os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, "-m", "pytest", os.path.join(base, "tests"), "-q", "-x" if False else
               "-rfEs", "-p", "no:cacheprovider"] + args + extra
        if kexpr:
            cmd += ["-k", kexpr]
        log = os.path.join(ROOT, "gpurun_out", f"ref_pytests_{name}.log")
        with open(log, "w") as f:
            f.write("$ " + " ".join(cmd) + "\n")
            f.flush()
            rc = subprocess.call(cmd, stdout=f, stderr=subprocess.STDOUT, env=env, cwd=base)
        tail = open(log).read().strip().splitlines()[-1:]
        print(name, "rc", rc, *tail)
        rc_all |= rc
    return rc_all


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["stage", "run"])
    ap.add_argument("--which", default="both", choices=["cufinufft", "finufft", "both"])
    ap.add_argument("-k", default=None)
    a, extra = ap.parse_known_args()
    if a.cmd == "stage":
        stage()
    else:
        sys.exit(run(a.which, a.k, extra))
