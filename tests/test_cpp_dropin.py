"""Drop-in check of the C++ header layer (hala_b200/gpu/): the reference's OWN templates and test bodies compiled on top of it.

CPU part (build container, where /root/reference exists): the unmodified hala::solve_cg template instantiates against the
replacement gpu/ directory (syntax-only compile) and the test binary builds.
GPU part: tests/_bin/ref_tests_on_b200 (built here by tests/cpp/Makefile, shipped with the snapshot) runs the reference's
cuda_core / cuda_blas1 / cuda_blas2(gemv) / cuda_sparse tests, its shared blas1/sparse test bodies through mixed_engine, and
solve_cg / solve_gmres on gpu_engine + mixed_engine against cpu_engine, for float, double and both complex types."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_bin", "ref_tests_on_b200")
HAVE_REF = os.path.isdir("/root/reference/wax")


@pytest.mark.skipif(not HAVE_REF, reason="reference headers only exist in the build container")
def test_reference_templates_compile_on_the_b200_layer():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "syntax"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not HAVE_REF, reason="reference headers only exist in the build container")
def test_dropin_binary_builds():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(BIN)


def test_header_layer_mentions_no_vendor_library():
    """No cuBLAS / cuSPARSE / cuSOLVER in the replacement layer."""
    import re
    for f in os.listdir(os.path.join(ROOT, "hala_b200", "gpu")):
        text = open(os.path.join(ROOT, "hala_b200", "gpu", f)).read()
        code = re.sub(r"//.*", "", text)
        assert not re.search(r"cublas[A-Z]|cusparse[A-Z]|cusolver[A-Z]|#include\s*<cublas|#include\s*<cusparse", code), f


@pytest.mark.gpu
def test_reference_tests_pass_on_the_b200_backend():
    if not os.path.exists(BIN):
        pytest.fail("tests/_bin/ref_tests_on_b200 missing: run __graft_entry__.build() in the build container first")
    env = dict(os.environ)
    import sysconfig            # the wheel's OpenBLAS (CPU side of the comparison) needs its bundled libgfortran
    blasdir = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
    env["LD_LIBRARY_PATH"] = blasdir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([BIN, "-v"], capture_output=True, text=True, timeout=300, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "Fail" not in out, out[-4000:]
    assert out.count("Pass") >= 40, out[-2000:]
