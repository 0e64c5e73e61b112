import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both tiers (CPU tier only loads it)."""
    lib = os.path.join(ROOT, "laris_b200", "liblaris_b200.so")
    if not os.path.isfile(lib):
        import __graft_entry__ as g
        g.build()
    return lib


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "tiny_reference.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def tiny():
    """The dataset the golden fixture was generated from (oracle/make_golden.py TINY)."""
    from laris_b200.synth import make_dataset
    return make_dataset(1, n=1500, G=300, G_u=160, P=120, C=5)
