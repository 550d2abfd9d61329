import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    # A `-m gpu` run on a box without a GPU must fail loudly, not silently pass on a fallback.
    pass


@pytest.fixture(scope="session")
def dev():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is available (there is no CPU fallback)")
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc, ~1 s).  The CUDA library is built by __graft_entry__.build(); if a checkout arrives
    without the git-ignored artefacts it is rebuilt here once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__

    __graft_entry__.ensure_built()
    import oracle

    oracle.build()
