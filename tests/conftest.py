import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def libfnx():
    """The product library (built in-tree if stale).  CPU tests only load it and look at symbols."""
    from fluidnexus_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import raster_oracle
    raster_oracle.build()
    return True


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
