import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def gpu(request):
    """Initialise libmf6gpu on cuda:0.  Under `-m gpu` (the GPU tier) a missing device is a hard failure -- the
    product path has no CPU fallback; in an unfiltered run on a machine without a usable driver the GPU tests
    are skipped instead of burying real CPU-side failures under init errors."""
    from modflow6_b200 import lib
    try:
        lib.init(0)
    except Exception as e:
        expr = request.config.getoption("-m") or ""
        if "gpu" in expr and "not gpu" not in expr or os.environ.get("MF6GPU_REQUIRE_GPU"):
            raise
        pytest.skip(f"no usable CUDA device: {e}")
    return lib
