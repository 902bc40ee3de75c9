import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the oracle (gcc) and libsdrgpu.so (nvcc cross-compiles without a GPU) once per session."""
    import oracle
    oracle.build()
    from sdrtrunk_b200 import build as product_build
    product_build.build()
    yield


@pytest.fixture(scope="session")
def gpu():
    from sdrtrunk_b200 import native
    native.init(0)
    return native
