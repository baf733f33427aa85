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


@pytest.fixture(params=["batch_kernels", "fused_kernel"])
def kernel_route(request, backend):
    """Small device batches take the one-launch fused kernel (marg_event_fused_kernel) by default; this fixture runs a
    test once on each route so that the warp-per-window batch kernels stay covered on the small fixtures too."""
    from is_vins_b200 import capi
    backend.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 0 if request.param == "batch_kernels" else 148)
    yield request.param
    backend.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 148)
