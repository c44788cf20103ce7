"""The CPU oracle against the reference's own known answers and against itself (numpy twin vs C).

The reference's tests pin only the Poisson term and the earth-norm bound
(tests/test_likelihood_correctness.jl:5-17, 164-182); everything else here is validation by
construction: two independent restatements (LAPACK eigh vs Jacobi) and physics known answers.
"""
import math

import numpy as np
import pytest

from oracle import oracle_c, oracle_np as O


# ---- known answers held by the reference's own tests ------------------------------------------------
def test_poisson_kat_from_reference_tests():
    e = np.array([10.0, 20.0, 30.0, 15.0]); m = np.array([12.0, 18.0, 29.0, 14.0])
    ref = sum((ei - mi) + mi * math.log(mi / ei) for ei, mi in zip(e, m))
    assert abs(O.poisson_log_likelihood(e, m) + ref) < 1e-10          # test_likelihood_correctness.jl:5-12
    assert abs(oracle_c.COracle.poisson(e, m)[0] + ref) < 1e-10
    n = np.array([1.0, 5.0, 10.0])
    assert abs(O.poisson_log_likelihood(n, n)) < 1e-12                # :14-17
    assert abs(oracle_c.COracle.poisson(n, n)[0]) < 1e-12
    with pytest.raises(ValueError):
        O.poisson_log_likelihood([1.0, 2.0], [1.0])                   # :19-20
    assert math.isfinite(O.poisson_log_likelihood([0.0, 1.0, 2.0], [0.0, 0.0, 3.0]))   # :26-29


def test_poisson_negative_rate_semantics():
    # likelihood_statistical.jl:8-15: warn above -1e-3, error at or below
    flags = []
    O.poisson_log_likelihood([-1e-4, 1.0], [1.0, 1.0], flags)
    assert flags == ["warn"]
    with pytest.raises(O.NegativeRateError):
        O.poisson_log_likelihood([-1.0, 2.0], [1.0, 2.0])
    _, fl = oracle_c.COracle.poisson(np.array([-1.0, 2.0]), np.array([1.0, 2.0]))
    assert fl & 4
    _, fl = oracle_c.COracle.poisson(np.array([-1e-4, 2.0]), np.array([1.0, 2.0]))
    assert fl == 2


def test_earth_norm_bounds_give_minus_inf(tables, asimov):
    # test_likelihood_correctness.jl:53-67,164-182
    p = np.array(tables["truth"], dtype=float)
    co = oracle_c.COracle(tables, asimov)
    for col, val in ((8, -0.01), (9, 2.01)):
        q = p.copy(); q[col] = val
        assert O.total_llh(q, tables, asimov) == -math.inf
        l, fl = co.eval_llh(q)
        assert l[0] == -math.inf and fl[0] == 1
    q = p.copy(); q[8], q[9] = 0.5, 1.9
    assert math.isfinite(O.total_llh(q, tables, asimov))


# ---- physics known answers ----------------------------------------------------------------------------------
def test_vacuum_limit_p1e(tables):
    t = dict(tables); t["earth_lookup"] = tables["earth_lookup"] * 0
    p = np.array(tables["truth"], dtype=float)
    o = O.compute_shared_osc_probs(p, t)
    want = (1 - 0.307) * (1 - 0.0224)       # c12^2 c13^2 = 0.6774768
    assert abs(want - 0.6774768) < 1e-12
    assert np.max(np.abs(o["p1e"] - want)) < 1e-13
    assert np.max(np.abs(o["nue_8B_night"] - o["nue_8B_day"][None, :])) < 1e-13


def test_pmns_first_element_and_unitarity():
    mp = O.OscPars(6e-5, math.asin(math.sqrt(0.307)), math.asin(math.sqrt(0.0224)))
    u = O.get_pmns(mp)
    assert abs(abs(u[0, 0]) ** 2 - (1 - 0.307) * (1 - 0.0224)) < 1e-15
    assert np.max(np.abs(u @ u.conj().T - np.eye(3))) < 1e-15
    assert u[0, 1].real > 0 and u[1, 0].real < 0    # U12[1,2] = +s12, U12[2,1] = -s12 (osc.jl:62)


def test_probabilities_are_probabilities(tables, draw):
    for p in draw(tables, 3, seed=11):
        o = O.compute_shared_osc_probs(p, tables)
        for k, v in o.items():
            assert np.all(v > -1e-12) and np.all(v < 1 + 1e-12), k
        assert np.allclose(o["nue_8B_day"] + o["nuother_8B_day"], 1.0)


# ---- restatement details ---------------------------------------------------------------------------------------
def test_block_average_shapes_and_errors():
    a = np.arange(12.0)
    assert np.array_equal(O.block_average(a, 2), np.array([0.5, 2.5, 4.5, 6.5, 8.5, 10.5]))
    m = np.arange(24.0).reshape(4, 6)
    assert O.block_average(m, (2, 2)).shape == (2, 3)
    assert O.block_average(m, (1, 2))[0, 0] == 0.5
    with pytest.raises(ValueError):
        O.block_average(a, 5)
    with pytest.raises(ValueError):
        O.block_average(m, (3, 2))


def test_all_es_backgrounds_share_norm_1(tables):
    # propagation_bg.jl:56-57: norm_index is never advanced for the ES components
    p = np.array(tables["truth"], dtype=float)
    base, _ = O.normalize_backgrounds(tables, O.view(p, tables))
    q = p.copy(); q[15] *= 3; q[16] *= 5          # ES_bg_norm_2, _3: sampled but unused
    same, _ = O.normalize_backgrounds(tables, O.view(q, tables))
    assert np.array_equal(base, same)
    q = p.copy(); q[14] *= 2
    dbl, _ = O.normalize_backgrounds(tables, O.view(q, tables))
    assert np.allclose(dbl, 2 * base)
    # CC norms do advance
    q = p.copy(); q[18] *= 2
    _, cc = O.normalize_backgrounds(tables, O.view(q, tables))
    _, cc0 = O.normalize_backgrounds(tables, O.view(p, tables))
    assert not np.allclose(cc, cc0) and np.allclose(cc - cc0, tables["raw_bg_cc"][1])


def test_asimov_llh_is_zero_and_perbin_mask(tables, asimov):
    p = np.array(tables["truth"], dtype=float)
    assert abs(O.total_llh(p, tables, asimov)) < 1e-6
    pb = O.perbin_llh(p, tables, asimov)
    k = int(tables["index_es"]) - 1
    assert np.all(pb["ES_day"][:k] == 0) and pb["ES_day"][-1] == 0      # below threshold and overflow bin
    assert np.all(pb["ES_night"][:, -1] == 0)


# ---- the two restatements agree --------------------------------------------------------------------------------
def test_c_and_numpy_oracles_agree(tables, asimov, draw):
    p = draw(tables, 6, wide=True)
    p[0] = tables["truth"]
    co = oracle_c.COracle(tables, asimov)
    osc, p1e = co.eval_oscprobs(p, want_p1e=True)
    rates = co.eval_rates(p)
    lnl, _ = co.eval_llh(p)
    for i, q in enumerate(p):
        on = O.compute_shared_osc_probs(q, tables)
        assert np.max(np.abs(p1e[i] - on["p1e"])) < 2e-12
        assert np.max(np.abs(osc[i, 0, 1:] - on["nue_8B_night"])) < 2e-12
        assert np.max(np.abs(osc[i, 1, 0] - on["nue_hep_day"])) < 1e-15
        rn = O.propagate_samples(q, tables)
        for k in rates:
            assert np.max(np.abs(rates[k][i] - rn[k]) / np.maximum(np.abs(rn[k]), 1e-300)) < 1e-11, k
        assert abs(2 * lnl[i] - 2 * O.total_llh(q, tables, asimov)) < 1e-6


def test_slow_mode_c_and_numpy_agree(tables):
    p = np.array(tables["truth"], dtype=float); p[0] = 0.33
    cs = oracle_c.COracle(tables, None, fast=False)
    osc = cs.eval_oscprobs(p)
    on = O.compute_shared_osc_probs(p, tables, fast=False)
    assert np.max(np.abs(osc[0, 0, 1:] - on["nue_8B_night"])) < 2e-12
    assert np.max(np.abs(osc[0, 0, 0] - on["nue_8B_day"])) < 1e-13
    # and the production-region integral really differs from the mean-density shortcut
    fast = O.compute_shared_osc_probs(p, tables, fast=True)
    assert np.max(np.abs(fast["nue_8B_day"] - on["nue_8B_day"])) > 1e-5
