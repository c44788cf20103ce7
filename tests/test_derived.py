"""Derived-mode reductions on the host against a literal per-sample loop (derive_variables_from_chain.jl:236-258)."""
import numpy as np

from oracle.derived_np import day_night_asymmetry, replay_block_sums, weighted_moments


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


def test_replay_block_sums_agree_with_the_vectorised_moments():
    rng = np.random.default_rng(2)
    B, nz, n = 9, 4, 6
    rates = {k: rng.uniform(0.5, 3, s) for k, s in (("ES_day", (B, n)), ("ES_night", (B, nz, n)), ("CC_day", (B, n)),
                                                    ("CC_night", (B, nz, n)), ("BG_ES_tot", (B, n)), ("BG_CC_tot", (B, n)))}
    perbin = {k: rng.normal(size=rates[k].shape) for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
    lref = {k: rng.normal(size=perbin[k].shape[1:]) for k in perbin}
    p = rng.uniform(0, 1, (B, 5)); w = rng.integers(1, 4, B).astype(float)
    out = replay_block_sums(rates, perbin, lref, p, w, 3, 2)
    m1, m2 = weighted_moments(rates["ES_night"], w, 3)
    assert np.allclose(out["mom_ES"][0][1:], m1) and np.allclose(out["mom_ES"][1][1:], m2)
    l = perbin["CC_night"] - lref["CC_night"]
    assert np.allclose(out["mom_CC"][2][1:], np.einsum("i,ijk->jk", w, l))
    assert np.allclose(out["mom_CC"][5][1:], np.einsum("i,ijk->jk", w * p[:, 2], l))
    assert np.allclose(out["ltot"][:, 2], l.sum(axis=(1, 2)))
    assert np.isclose(out["scalars"][0], w.sum()) and np.isclose(out["scalars"][5 + 8 + 1], (w * l.sum(axis=(1, 2)) ** 2).sum())
    assert np.allclose(out["bg_mom_ES"][1], (w[:, None] * (0.5 * rates["BG_ES_tot"]) ** 2).sum(axis=0))
