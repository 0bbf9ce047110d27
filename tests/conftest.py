import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def gb():
    """The product library; builds it if the .so is missing (nvcc cross-compiles on CPU)."""
    import gamut_b200
    from gamut_b200 import _lib
    if not os.path.exists(_lib.LIBPATH):
        from gamut_b200 import build
        build.build()
    _lib.lib()
    return gamut_b200
