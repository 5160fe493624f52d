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
