import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _library_matches_sources():
    """The engine library is git-ignored and travels with the tree: rebuild it (nvcc cross-compiles without a GPU) when the
    hash compiled into it differs from the hash of the sources, so that no test ever runs a stale binary."""
    from m2trans_b200 import build
    if build._stale():
        build.build()
