import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


# ---------------------------------------------------------------------------- synthetic inputs
def gaussian_image(N, amp=10.0):
    """scripts/demo_iso_sn_voronoi.py:20-26 scaled to N (sigma = 50*N/256); SURVEY.md 8d."""
    s = 50.0 * N / 256.0
    x, y = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    return amp * np.exp(-((x - N / 2.0) ** 2.0 / (2.0 * s ** 2.0) + (y - N / 2.0) ** 2.0 / (2.0 * s ** 2.0)))


def sampled_seeds(w, k, seed):
    """SURVEY.md 8d seed recipe: k pixels drawn with p ~ w, jittered by U(-.5,.5) (tie-free)."""
    rng = np.random.default_rng(seed)
    H, W = w.shape
    p = w.ravel() / w.sum()
    sel = rng.choice(H * W, k, replace=False, p=p)
    gx = (sel % W).astype(float)
    gy = (sel // W).astype(float)
    return np.column_stack((gx, gy)) + rng.uniform(-0.5, 0.5, (k, 2))


@pytest.fixture(scope="session")
def engine_cls():
    """The real GPU engine; fails (not skips) when the CUDA library is missing on a GPU box."""
    from tess_b200.engine import Engine
    return Engine
