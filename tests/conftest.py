import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference recorded by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs.npz"))


@pytest.fixture(scope="session")
def golden_f1():
    """Triangular-solve / ILU(0) outputs of the unmodified reference (tests/golden/make_golden.py f1)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_f1.npz"))


@pytest.fixture(scope="session")
def golden_f2():
    """SpMM / batch-CG outputs of the unmodified reference (tests/golden/make_golden.py f2)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_f2.npz"))


@pytest.fixture(scope="session")
def ref_tests():
    """Vectors harvested from the reference's own tests (file:line inside the JSON)."""
    with open(os.path.join(ROOT, "tests", "golden", "ref_tests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    binding.build(with_ref=False)
    return binding.oracle()


@pytest.fixture(scope="session")
def engine():
    import hala_b200 as hb
    if hb.gpu_device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    return hb.gpu_engine(0)
