import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu():
    """Initialise the native library on cuda:0. GPU tests must FAIL (not skip) without a device:
    there is no CPU fallback to hide behind."""
    from basic_sparse_matrix_b200 import gpu as g
    g.init(0)
    return g
