"""GPU parity on the other BASELINE.json configurations.

C2 (the LLH scan) is bench.py's workload; these are the parity-test cases.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def engine(tables, asimov):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    with SolarOscEngine(tables, data=asimov) as e:
        yield e


# ---- C1: single point at truth + the literal tests/benchmark_likelihoods.jl shapes --------------------------
def test_c1_single_point_at_truth(engine, tables):
    from oracle import oracle_c
    got = engine.eval_llh(np.asarray(tables["truth"], dtype=float))
    assert got.shape == (1,) and abs(got[0]) < 1e-6


def test_c1_fake_200x200_likelihood_shape(engine):
    """benchmark_likelihoods.jl:23-74: 200 ES + 200 CC bins x (1 day + 200 zenith rows) = 80 400 Poisson terms,
    rates filled with one scale, data U{10..50}."""
    from oracle import oracle_c
    rng = np.random.default_rng(17)
    m = rng.integers(10, 51, (4, 80400)).astype(float)
    e = np.repeat(np.array([[1.0], [17.5], [30.0], [49.0]]), 80400, axis=1)
    got = engine.poisson_llh(e, m)
    ref = np.array([oracle_c.COracle.poisson(e[i], m[i])[0] for i in range(4)])
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12


# ---- C3: MCMC-style batches of independent prior draws ------------------------------------------------------------
@pytest.mark.parametrize("batch", [10, 4096])
def test_c3_prior_draw_batches(engine, tables, asimov, draw, batch):
    from oracle import oracle_c
    p = draw(tables, batch, seed=1000 + batch)
    ref, rfl = oracle_c.COracle(tables, asimov).eval_llh(p)
    got, fl = engine.eval_llh(p, return_flags=True)
    assert np.array_equal(fl, rfl)
    assert np.max(np.abs(2 * got - 2 * ref)) < 1e-6


# ---- C4: Earth-density systematics on dense zenith x energy grids ---------------------------------------------------
def _dense(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def _earth_draws(t, n, seed):
    rng = np.random.default_rng(seed)
    p = np.tile(np.asarray(t["truth"], dtype=float), (n, 1))
    p[:, 8:14] = np.clip(rng.normal(1.0, 0.12, (n, 6)), 0.0, 2.0)      # truncN(1, 0.12; 0, 2)
    p[0, 0], p[0, 2] = 0.4, 2e-4                                       # checkOscGridConvergence.jl:40-42 worst case
    return p


def _oracle_data(t):
    """Observed data of a dense configuration: the C oracle's rates at the truth (an input table like any other)."""
    from oracle import oracle_c
    r = oracle_c.COracle(t, None).eval_rates(np.asarray(t["truth"], dtype=float))
    return {k: r[k][0] for k in ("ES_day", "ES_night", "CC_day", "CC_night")}


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_c4_dense_200x200_matches_oracle():
    """200 zenith x 200 energies with the full detector tables: oscillation stage, rates and the likelihood
    (ES+CC, CC alone, night-only CC) through the row-chunked fold, against the oracle."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_c
    t = _dense("tables_dense200.npz")
    p = _earth_draws(t, 3, 5)
    ref_osc, ref_p1e = oracle_c.COracle(t, None).eval_oscprobs(p, want_p1e=True)
    data = _oracle_data(t)
    with SolarOscEngine(t, data=data) as e:
        p1e = e.eval_p1e(p)
        osc = e.eval_oscprobs(p)
        lnl, fl = e.eval_llh(p, return_flags=True)
        rates = e.eval_rates(p)
        perbin = e.eval_perbin(p[:1])
        big = _earth_draws(t, 300, 6)                                  # more points than SMs
        lnl_big = e.eval_llh(big)
    with SolarOscEngine(t, data=data) as e2:                           # a fresh handle whose FIRST call is a graph-replayed small
        lnl_200 = e2.eval_llh(big[:200])                               # batch of more points than SMs (workspace allocated before the capture)
    assert np.array_equal(lnl_200, lnl_big[:200])
    assert p1e.shape == (3, 200, 200)
    assert np.max(np.abs(p1e - ref_p1e)) < 5e-12
    assert np.max(np.abs(osc["nue_8B_night"] - ref_osc[:, 0, 1:])) < 5e-12
    assert np.max(np.abs(osc["nue_hep_day"] - ref_osc[:, 1, 0])) < 1e-14
    orc = oracle_c.COracle(t, data)
    ref_lnl, ref_fl = orc.eval_llh(p)
    ref_rates = orc.eval_rates(p)
    assert np.array_equal(fl, ref_fl)
    assert np.max(np.abs(2 * lnl - 2 * ref_lnl)) < 1e-6                 # BASELINE.json tolerance
    for k in ("ES_day", "ES_night", "CC_day", "CC_night"):
        assert rates[k].shape == ref_rates[k].shape
        assert _rel(rates[k], ref_rates[k]) < 1e-9, k
    assert perbin["CC_night"].shape == (1, 200, 30) and np.all(perbin["CC_night"][:, :, -1] == 0)
    assert np.max(np.abs(2 * lnl_big - 2 * orc.eval_llh(big)[0])) < 1e-6
    # night-only CC (BASELINE config 4): ES off, the day vector out of the sum
    ref_n, _ = oracle_c.COracle(t, data, es_mode=False, night_only=True).eval_llh(p)
    with SolarOscEngine(t, data=data, es_mode=False, night_only=True) as e:
        got_n = e.eval_llh(p)
    assert np.max(np.abs(2 * got_n - 2 * ref_n)) < 1e-6
    assert np.max(np.abs(ref_n - ref_lnl)) > 1e-3                       # (it is a different number)


def test_c4_dense_1000x1000_sweep():
    """The 1000 cos z x 1000 E grid of utils/checkOscGridConvergence.jl:36-46: the probability tile (8 MB per point)
    is 35 x one SM's shared memory, the fold runs row-chunked from HBM -- oscillation stage, CC night rates and the
    night-only CC likelihood against the oracle."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_c
    t = _dense("tables_dense1000.npz")
    p = _earth_draws(t, 2, 9)
    ref_osc, ref_p1e = oracle_c.COracle(t, None).eval_oscprobs(p[:1], want_p1e=True)
    data = _oracle_data(t)
    orc = oracle_c.COracle(t, data, es_mode=False, night_only=True)
    ref_lnl, ref_fl = orc.eval_llh(p)
    ref_rates = orc.eval_rates(p[:1])
    with SolarOscEngine(t, data=data, es_mode=False, night_only=True) as e:
        p1e = e.eval_p1e(p)
        assert p1e.shape == (2, 1000, 1000)
        assert np.max(np.abs(p1e[0] - ref_p1e[0])) < 5e-12
        osc = e.eval_oscprobs(p[:1])
        assert np.max(np.abs(osc["nue_8B_night"][0] - ref_osc[0, 0, 1:])) < 5e-12
        v = p[1].copy(); v[8:14] = 0.0                                  # vacuum: size-independent property
        assert np.max(np.abs(e.eval_p1e(v) - (1 - v[0]) * (1 - v[1]))) < 1e-14
        lnl, fl = e.eval_llh(p, return_flags=True)
        rates = e.eval_rates(p[:1])
        oob = p.copy(); oob[1, 9] = 2.5
        lo, flo = e.eval_llh(oob, return_flags=True)
    assert np.array_equal(fl, ref_fl) and np.max(np.abs(2 * lnl - 2 * ref_lnl)) < 1e-6
    assert _rel(rates["CC_night"], ref_rates["CC_night"]) < 1e-9 and _rel(rates["CC_day"], ref_rates["CC_day"]) < 1e-9
    assert np.all(rates["ES_night"] == 0)                               # propagation_reco.jl:12-15
    assert lo[0] == lnl[0] and lo[1] == -np.inf and flo[1] == 1         # earth-norm bound (likelihood_debug.jl:1-4)


def test_c4_night_only_cc_default_grid(tables, draw):
    """Night-only CC likelihood with earth-norm draws, B = 4096, on the default grid: ES off (CC day row still in the
    sum), then truly night-only (solarosc_config.night_only: the day vector out of the sum), each against the oracle
    configured the same way; 4096 points take the fused kernel, 16 the small-batch kernels."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_c, oracle_np
    data = oracle_np.asimov_data(tables)
    p = np.tile(np.asarray(tables["truth"], dtype=float), (4096, 1))
    p[:, 8:14] = np.clip(np.random.default_rng(3).normal(1.0, 0.12, (4096, 6)), 0.0, 2.0)
    p[:64, 0] = 0.33                                                   # off the Asimov point: the day term is not zero
    ref, _ = oracle_c.COracle(tables, data, es_mode=False).eval_llh(p)
    with SolarOscEngine(tables, data=data, es_mode=False) as e:
        got = e.eval_llh(p)
        r = e.eval_rates(p[:2])
    assert np.max(np.abs(2 * got - 2 * ref)) < 1e-6
    assert np.all(r["ES_day"] == 0) and np.all(r["ES_night"] == 0)     # propagation_reco.jl:12-15
    ref_n, _ = oracle_c.COracle(tables, data, es_mode=False, night_only=True).eval_llh(p)
    with SolarOscEngine(tables, data=data, es_mode=False, night_only=True) as e:
        got_n, got_small, got_grp = e.eval_llh(p), e.eval_llh(p[:16]), e.eval_llh(p, group=1)
        r_n = e.eval_rates(p[:2])
    assert np.max(np.abs(2 * got_n - 2 * ref_n)) < 1e-6 and np.max(np.abs(2 * got_small - 2 * ref_n[:16])) < 1e-6
    assert np.max(np.abs(ref_n - ref)) > 1e-3
    assert np.array_equal(r_n["CC_day"], r["CC_day"])                    # rates do not change, only the sum


# ---- C5: derived quantities over a stored chain ---------------------------------------------------------------------------
def test_c5_day_night_asymmetry_and_moments(engine, tables, asimov, draw):
    """derive_variables_from_chain.jl:231-258,285-444: asymmetries from thresholded sums, weighted rate moments."""
    from oracle.derived_np import day_night_asymmetry, weighted_moments
    from oracle import oracle_c
    p = draw(tables, 64, seed=77)
    w = np.random.default_rng(1).integers(1, 5, 64).astype(float)
    got = engine.eval_rates(p)
    ref = oracle_c.COracle(tables, asimov).eval_rates(p)
    for ch, idx in (("ES", int(tables["index_es"])), ("CC", int(tables["index_cc"]))):
        a = day_night_asymmetry(got[ch + "_day"], got[ch + "_night"], got["BG_" + ch + "_tot"], idx)
        b = day_night_asymmetry(ref[ch + "_day"], ref[ch + "_night"], ref["BG_" + ch + "_tot"], idx)
        assert np.max(np.abs(a - b)) < 1e-9
        m1, m2 = weighted_moments(got[ch + "_night"], w, idx)
        r1, r2 = weighted_moments(ref[ch + "_night"], w, idx)
        assert np.max(np.abs(m1 - r1) / np.maximum(np.abs(r1), 1e-300)) < 1e-9
        assert np.max(np.abs(m2 - r2) / np.maximum(np.abs(r2), 1e-300)) < 1e-9
    # the same asymmetries reduced on the device (16 bytes per sample instead of the rate arrays)
    a_es, a_cc = engine.eval_asymmetry(p)
    assert np.max(np.abs(a_es - day_night_asymmetry(ref["ES_day"], ref["ES_night"], ref["BG_ES_tot"], int(tables["index_es"])))) < 1e-9
    assert np.max(np.abs(a_cc - day_night_asymmetry(ref["CC_day"], ref["CC_night"], ref["BG_CC_tot"], int(tables["index_cc"])))) < 1e-9
    big = draw(tables, 600, seed=78)          # large enough for the fused kernel
    b_es, b_cc = engine.eval_asymmetry(big)
    rb = engine.eval_rates(big)
    assert np.max(np.abs(b_cc - day_night_asymmetry(rb["CC_day"], rb["CC_night"], rb["BG_CC_tot"], int(tables["index_cc"])))) < 1e-10
    assert np.max(np.abs(b_es - day_night_asymmetry(rb["ES_day"], rb["ES_night"], rb["BG_ES_tot"], int(tables["index_es"])))) < 1e-10  # background-dominated: rounding order shows
    pb = engine.eval_perbin(p[:4])
    assert pb["CC_night"].shape == (4, 50, 30) and np.all(pb["CC_night"][:, :, -1] == 0)


def test_c5_device_block_sums_match_the_replay_loop(engine, tables, asimov, draw):
    """solarosc_eval_derived against the replay loop (derive_variables_from_chain.jl:260-444) run in numpy on the
    oracle's rates and per-bin terms; 1100 samples span two internal chunks."""
    from oracle.derived_np import replay_block_sums
    from oracle import oracle_c, oracle_np
    n = 1100
    p = draw(tables, n, seed=79)
    w = np.random.default_rng(2).integers(1, 6, n).astype(float)
    truth = np.asarray(tables["truth"], dtype=float)[None, :]
    lref = {k: v[0] for k, v in engine.eval_perbin(truth).items()}
    got = engine.eval_derived(p, w, lref)
    orc = oracle_c.COracle(tables, asimov)
    rates = orc.eval_rates(p)
    data = asimov
    perbin = {k: np.empty_like(rates[k]) for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
    for ch, idx in (("ES", int(tables["index_es"])), ("CC", int(tables["index_cc"]))):
        obs = np.vstack([np.asarray(data[ch + "_day"])[None, :], np.asarray(data[ch + "_night"])])
        for i in range(n):
            r = np.vstack([rates[ch + "_day"][i][None, :], rates[ch + "_night"][i]])
            t = np.zeros_like(r)
            t[:, idx - 1:-1] = np.stack([oracle_np.perbin_poisson_log_likelihood(r[z, idx - 1:-1], obs[z, idx - 1:-1]) for z in range(len(r))])
            perbin[ch + "_day"][i], perbin[ch + "_night"][i] = t[0], t[1:]
    ref = replay_block_sums(rates, perbin, lref, p, w, int(tables["index_es"]), int(tables["index_cc"]))

    def close(a, b, rel):
        scale = np.maximum(np.abs(b), 1e-12 * np.max(np.abs(b)) + 1e-300)
        return np.max(np.abs(a - b) / scale) < rel

    for ch in ("ES", "CC"):
        assert np.max(np.abs(got[ch + "_day"] - rates[ch + "_day"]) / np.maximum(np.abs(rates[ch + "_day"]), 1e-300)) < 1e-9
        for q in range(6):
            assert close(got["mom_" + ch][q], ref["mom_" + ch][q], 1e-8 if q >= 2 else 1e-9), (ch, q)
        assert np.all(got["mom_" + ch][:2, :, :int(tables["index_" + ch.lower()]) - 1] == 0)      # below threshold
        assert close(got["bg_mom_" + ch], ref["bg_mom_" + ch], 1e-12)
    assert np.max(np.abs(got["ltot"] - ref["ltot"])) < 1e-6
    assert close(got["scalars"], ref["scalars"], 1e-9)
    # block sums add: two calls on the halves equal one call on the whole
    a, b = engine.eval_derived(p[:500], w[:500], lref), engine.eval_derived(p[500:], w[500:], lref)
    assert close(a["mom_CC"] + b["mom_CC"], got["mom_CC"], 1e-12) and close(a["scalars"] + b["scalars"], got["scalars"], 1e-12)


# ---- C2: the full 512 x 512 x 27 LLH scan through size-independent properties --------------------------------------------
def test_c2_full_scan_properties(engine, tables, asimov):
    """BASELINE.json configs[1] at full size (7.08 M points, llhScan.jl:35-37,48-108).  The oracle cannot visit every
    point in test time, so: (1) an evenly spaced sample against the C oracle at the stated tolerance; (2) the scan's
    structure -- the centre points of a cell row reproduce a direct evaluation of those points in a small batch,
    bit for bit, whatever their position in the 7 M-point call; (3) a shuffled copy of a slice of the scan returns
    the same values, permuted; (4) the grouped evaluation (27 points of a cell share the oscillation stage) agrees
    with the ungrouped one to summation order; (5) the minimum of -2lnL over the scan sits in the cell nearest the truth."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import scan_params
    from oracle import oracle_c
    grid = 512
    p, n_cell = scan_params(tables, grid, float(tables["truth"][1]))
    assert p.shape == (grid * grid * 27, 19) and n_cell == 27
    lnl, fl = engine.eval_llh(p, return_flags=True)
    assert np.all(np.isfinite(lnl)) and not np.any(fl)
    # (1) sample against the oracle
    idx = np.arange(0, len(p), len(p) // 300)[:300]
    ref, _ = oracle_c.COracle(tables, asimov).eval_llh(np.ascontiguousarray(p[idx]))
    assert np.max(np.abs(2 * lnl[idx] - 2 * ref)) < 1e-6
    # (2) position independence, bit for bit
    row = (np.arange(grid) * grid + 200) * 27 + 26            # centre point (no nuisance shift) of the cells of column 200
    assert np.array_equal(engine.eval_llh(np.ascontiguousarray(p[row])), lnl[row])
    # (3) permutation
    rng = np.random.default_rng(5)
    sl = slice(3_000_000, 3_200_000)
    perm = rng.permutation(200_000)
    assert np.array_equal(engine.eval_llh(np.ascontiguousarray(p[sl][perm])), lnl[sl][perm])
    # (4) grouped evaluation
    g = engine.eval_llh(p, group=27)
    assert np.max(np.abs(g - lnl)) < 1e-9
    # (5) the scan's minimum is at the truth (Asimov data): sin2_th12 = 0.307, dm2_21 = 6e-5
    centre = lnl.reshape(grid, grid, 27)[:, :, 26]
    i, j = np.unravel_index(np.argmax(centre), centre.shape)
    v12, vdm = np.linspace(0.2, 0.4, grid), np.linspace(4e-5, 1.5e-4, grid)
    assert abs(v12[i] - 0.307) <= (v12[1] - v12[0]) and abs(vdm[j] - 6e-5) <= (vdm[1] - vdm[0])
