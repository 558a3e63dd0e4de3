import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the GPU tests; fails loudly (no CPU fallback) without a device."""
    import fflins
    c = fflins.Context(fflins.default_config())
    yield c
    c.close()


def new_ctx(**kw):
    import fflins
    return fflins.Context(fflins.default_config(**kw))
