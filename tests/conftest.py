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
    """The in-tree C-ABI library, rebuilt when a source is newer than it (nvcc cross-compiles without a GPU; the
    mtime checks make an up-to-date build a no-op), so a stale library is never tested after an edit.  A developer
    override (MINIAERO_B200_LIB, tools/build_variants.py) is used as is."""
    from miniaero_b200 import _abi, build
    if not os.environ.get("MINIAERO_B200_LIB"):
        try:
            build.build()
        except RuntimeError:
            if not os.path.isfile(_abi.LIB_PATH):   # no nvcc on this box: the prebuilt library that travelled is used
                raise
    return _abi.load()
