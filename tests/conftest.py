import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def tables():
    return dict(np.load(os.path.join(GOLDEN, "tables_escc.npz")))


@pytest.fixture(scope="session")
def asimov(tables):
    from oracle import oracle_np
    return oracle_np.asimov_data(tables)


def draw_params(tables, n, seed=20261019, wide=False):
    """Seeded prior draws (config.yaml:38-52; earth sigma 0.12; BG norms N(norm, 0.02 norm))."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    truth = np.asarray(tables["truth"], dtype=np.float64)
    p = np.tile(truth, (n, 1))
    nl = len(tables["earth_lookup"])
    p[:, 0] = rng.uniform(0.15, 0.5, n)
    p[:, 1] = np.clip(rng.normal(0.022, 0.0007, n), 0.005, 0.035)
    p[:, 2] = rng.uniform(2e-5, 3e-4, n)
    p[:, 3] = np.clip(rng.normal(5.25e6, 0.28e6, n), 1.5e6, 1e7)
    p[:, 4] = rng.uniform(0, 2e4, n)
    p[:, 5] = np.clip(rng.normal(1.0, 0.1, n), 0.1, 2.0)
    p[:, 6] = rng.normal(0.0, 0.1, n)
    p[:, 7] = rng.normal(0.0, 0.05, n)
    p[:, 8:8 + nl] = np.clip(rng.normal(1.0, 0.3 if wide else 0.12, (n, nl)), 0.0, 2.0)
    k = 8 + nl
    p[:, k:] = truth[k:] * rng.normal(1.0, 0.02, (n, p.shape[1] - k))
    return p


@pytest.fixture(scope="session")
def draw():
    return draw_params
