import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device; run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (cross-compiled here, prebuilt on the GPU box)."""
    from denseretrievaltoolkits_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()
