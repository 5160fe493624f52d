"""The reference's own C / CUDA API contract programs, linked against OUR library.

oracle/build.py::build_ref_api_tests compiles test/cuda/{test_makeplan.c, public_api_test.c,
cufinufft_error_handling.cu, cufinufft_multigpu_test.cu, cufinufft_simple_test.cu} from the
sources where they lie under /root/reference (unmodified; only the link target changes from
libcufinufft to finufft_b200/libfinufft_b200.so) into oracle/_ref/bin/, which travels to the GPU
box.  Each program is its own judge: exit code 0 = pass (77 = skipped: the multi-GPU program
needs two devices), exactly as the reference's CMake registers them (test/cuda/CMakeLists.txt).
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
PROGRAMS = ["test_makeplan", "public_api_test", "cufinufft_error_handling",
            "cufinufft_simple_test", "cufinufft_multigpu_test"]


@pytest.mark.parametrize("prog", PROGRAMS)
def test_reference_api_program(cuda, prog):
    path = os.path.join(BIN, prog)
    if not os.path.exists(path):
        from oracle import build as ob
        ob.build_ref_api_tests()
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "finufft_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([path], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    tail = (r.stdout[-1500:] + "\n" + r.stderr[-1500:]).strip()
    if r.returncode == 77:
        pytest.skip(f"{prog}: skipped by the program itself (needs 2 GPUs)\n{tail}")
    assert r.returncode == 0, f"{prog} exited with {r.returncode}\n{tail}"


@pytest.mark.parametrize("pkg", ["cufinufft", "finufft"])
def test_reference_python_suites(cuda, pkg):
    """The reference's own pytest suites (python/cufinufft/tests, python/finufft/test: 2927 + 659
    cases, unmodified) with its ctypes bindings loading THIS library under the reference's
    library file name.  tools/ref_pytests.py stages the packages into oracle/_ref/py/ in the
    container that holds /root/reference; they travel with the snapshot."""
    import sys
    staged = os.path.join(ROOT, "oracle", "_ref", "py", pkg)
    if not os.path.isdir(staged):
        if os.path.isdir("/root/reference/python"):
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ref_pytests.py"),
                                   "stage"])
        else:
            pytest.skip("oracle/_ref/py not staged (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_pytests.py"), "run",
                        "--which", pkg], capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert " passed" in r.stdout and "failed" not in r.stdout
