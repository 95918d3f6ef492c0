import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def handle():
    from gp_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "gp_derivs_golden.npz"))


@pytest.fixture(scope="session")
def ch2_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ch2_golden.npz"))


@pytest.fixture(scope="session")
def westbrook():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "westbrook_xy.npz"))
