import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def product_lib():
    """The sm_100a product library.  Loading it needs no GPU (host helpers work on CPU)."""
    from lapx_b200 import api, build
    build.build_product()
    return api.load_product()


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle — test infrastructure only."""
    from lapx_b200 import api, build
    return api.load_library(build.build_oracle())
