import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def dev():
    import torch
    return torch.device("cuda:0")


def random_boxes(rng, n, min_size=0.01, max_size=0.3, clusters=0):
    """Normalised (y1,x1,y2,x2) boxes; clusters>0 makes heavily overlapping groups."""
    if clusters:
        centres = rng.uniform(0.15, 0.85, (clusters, 2))
        sizes = rng.uniform(min_size * 2, max_size, (clusters, 2))
        which = rng.integers(0, clusters, n)
        c = centres[which] + rng.normal(0, 0.01, (n, 2))
        s = sizes[which] * np.exp(rng.normal(0, 0.08, (n, 2)))
    else:
        c = rng.uniform(0, 1, (n, 2))
        s = rng.uniform(min_size, max_size, (n, 2))
    b = np.concatenate([c - s / 2, c + s / 2], axis=1)
    return np.clip(b, 0, 1).astype(np.float32)
