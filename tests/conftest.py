import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from __graft_entry__ import load_oracle, load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def orc():
    return load_oracle()


@pytest.fixture(scope="session")
def native_lib(pkg):
    """The C-ABI library, built if necessary (nvcc cross-compiles without a GPU)."""
    if not os.path.exists(pkg.library_path()):
        pkg.build_native()
    return pkg.load_library()
