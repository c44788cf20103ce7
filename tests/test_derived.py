"""Derived-mode reductions on the host against a literal per-sample loop (derive_variables_from_chain.jl:236-258)."""
import numpy as np

from dune_solaroscfitter_b200.derived import day_night_asymmetry, weighted_moments


def test_asymmetry_matches_the_per_sample_loop():
    rng = np.random.default_rng(0)
    day = rng.uniform(1, 10, (5, 6)); night = rng.uniform(0, 1, (5, 4, 6)); bg = rng.uniform(0, 2, (5, 6))
    idx = 3
    got = day_night_asymmetry(day, night, bg, idx)
    for i in range(5):
        b = bg[i, idx - 1:].sum()
        n = night[i][:, idx - 1:].sum() - 0.5 * b
        d = day[i, idx - 1:].sum() - 0.5 * b
        assert abs(got[i] - 2 * (d - n) / (d + n)) < 1e-14
    z = day_night_asymmetry(np.zeros((1, 6)), np.zeros((1, 4, 6)), np.zeros((1, 6)), idx)
    assert z[0] == 0.0            # zero denominator guard


def test_weighted_moments_leave_sub_threshold_bins_zero():
    rng = np.random.default_rng(1)
    r = rng.uniform(0, 3, (7, 4, 6)); w = rng.integers(1, 4, 7).astype(float)
    m1, m2 = weighted_moments(r, w, 3)
    assert np.all(m1[:, :2] == 0) and np.all(m2[:, :2] == 0)
    assert np.allclose(m1[:, 2:], np.einsum("i,ijk->jk", w, r[:, :, 2:]))
    assert np.allclose(m2[:, 2:], np.einsum("i,ijk->jk", w, r[:, :, 2:] ** 2))
