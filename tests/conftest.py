import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The reference package (build container only; absent on the GPU box) must be importable BEFORE
# artensor_b200 is imported: artensor_b200.TensorNetworkSimulation subclasses the reference's class
# when it can.  Tests that need the reference skip themselves when it is missing.
_REF = os.environ.get("ARTENSOR_REFERENCE", "/root/reference")
if os.path.isdir(os.path.join(_REF, "artensor")) and _REF not in sys.path:
    sys.path.append(_REF)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import numpy as np
    from artensor_b200.cases import load_case
    case = load_case(os.path.join(GOLDEN, f"{name}.case.gz"))
    exp = np.load(os.path.join(GOLDEN, f"{name}.expected.npz"))
    return case, exp
