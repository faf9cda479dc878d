import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_cases(fname):
    """tests/golden/<fname> -> {case: {key: array}} (keys are 'case/key' in the npz)."""
    z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    cases = {}
    for k in z.files:
        if "/" not in k:
            continue
        c, kk = k.split("/", 1)
        cases.setdefault(c, {})[kk] = z[k]
    return cases


@pytest.fixture(scope="session")
def golden():
    return load_cases
