import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (CANON restatement + REF runner). Test infrastructure only."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def dbt():
    """The product: ctypes face of libdbt_b200.so (built in-tree; no fallback)."""
    mod = importlib.import_module("database-technology-algorithms_b200")
    if not os.path.exists(mod.LIB_PATH):
        build = importlib.import_module("database-technology-algorithms_b200.build")
        build.build()
    mod.lib()
    return mod
