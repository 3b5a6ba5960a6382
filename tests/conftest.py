import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "frontend.npz"))


@pytest.fixture(scope="session")
def golden_banks():
    return np.load(os.path.join(GOLDEN, "filterbanks.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import frontend_oracle
    return frontend_oracle


@pytest.fixture(scope="session")
def clips(oracle, golden):
    x = oracle.synth_clips(golden["clip_indices"])
    assert float(np.sum(x, dtype=np.float64)) == float(golden["clips_checksum"][0]), \
        "synthetic generator drifted from the one the golden vectors were made with"
    return x


@pytest.fixture(autouse=True)
def _deterministic_draws(request):
    """Every test starts from the same torch seed (CPU and CUDA): a test that draws without its own generator sees the same
    numbers in every run and in every order, so a pass here is a pass on the driver's box."""
    import zlib

    import torch
    torch.manual_seed(zlib.crc32(request.node.name.encode()))
    yield
