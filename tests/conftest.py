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


def branch_tables(tables, inclusive=0, angular=0):
    """Measurement tables switched to another analysis branch (VD block of config.yaml:123-156): the extra
    tables of tests/golden/tables_ext.npz merged in, the ES efficiencies boosted as response.jl:147-160 does for
    the inclusive modes."""
    ext = dict(np.load(os.path.join(GOLDEN, "tables_ext.npz")))
    t = dict(tables)
    t.update({k: ext[k] for k in ("n_cos", "resp_cc_incl", "eff_cc_incl", "resp_es_angular", "resp_bg_angular", "angular_mask",
                                  "resp_cc_misid_nue", "resp_cc_misid_nuother", "cc_f_above", "cc_f_below")})
    if inclusive:
        t["eff_es_nue"], t["eff_es_nuother"] = ext["eff_es_nue_boosted"], ext["eff_es_nuother_boosted"]
    t["inclusive"], t["angular"] = np.int32(inclusive), np.int32(angular)
    return t


@pytest.fixture(scope="session")
def branches():
    return branch_tables


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
