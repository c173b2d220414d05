import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def gpu():
    """Initialise libmf6gpu on cuda:0; fails loudly when there is no GPU."""
    from modflow6_b200 import lib
    lib.init(0)
    return lib
