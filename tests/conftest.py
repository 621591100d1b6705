import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the C oracle are build artefacts: make sure both exist."""
    from mindtheedge_b200 import build
    build.build()
    from oracle import pr_counts
    pr_counts._lib()


def load_cases(npz_name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN, npz_name))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    return cases
