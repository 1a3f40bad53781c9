import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def yeast_pyramid():
    from graal_b200.level import yeast_shaped_pyramid
    return yeast_shaped_pyramid()


@pytest.fixture(scope="session")
def small_pyramid():
    """A small pyramid for fast CPU tests: 6 contigs, 600 level-0 frags, 3 levels."""
    from graal_b200.level import build_synthetic_pyramid
    return build_synthetic_pyramid([300_000, 200_000, 150_000, 90_000, 40_000, 6_000], 600, 3,
                                   seed=11, cis_rowsum=300.0, v_inter=0.05)
