import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The in-tree C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from miniaero_b200 import _abi, build
    if not os.path.isfile(_abi.LIB_PATH):
        build.build()
    return _abi.load()
