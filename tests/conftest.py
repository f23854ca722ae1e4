import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def windows_set():
    """The four window sets of the reference's tests/test_paramgen.py:4-28."""
    import numpy as np
    return [
        [(0, 0, np.array([1.0]))],
        [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5]))],
        [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0]))],
        [(0, 0, np.array([1.0])), (2, 2, np.array([1.0, -8.0, 0.0, 8.0, -1.0]) / 12.0),
         (2, 2, np.array([-1.0, 16.0, -30.0, 16.0, -1.0]) / 12.0)],
    ]


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "mlpg_reference_golden.npz"))


@pytest.fixture(scope="session")
def dtw_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "dtw_restated_golden.npz"))


def rel_err(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
