import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def backend():
    """One MargBackend on cuda:0.  Fails loudly (no CPU fallback) if the library or GPU is missing."""
    from is_vins_b200 import MargBackend
    be = MargBackend(0)
    yield be
    be.close()
