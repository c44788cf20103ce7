"""Binnings other than config.yaml's defaults: odd numbers of true-energy, reco-energy and zenith bins, fewer Earth
layers, several sub-paths per zenith bin (N_COSZ_SUB > 1), missing nuisance columns -- synthetic tables, GPU against
the numpy oracle.  Exercises the padding of the fold's tiles (K not a multiple of 4, rows not a multiple of 8, reco
bins not a multiple of 8) and the generic path-table code."""
import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def synthetic_tables(seed, n_etrue, n_cosz, n_sub, n_layers, n_es, n_cc, idx_es, idx_cc, n_bg_es=2, n_bg_cc=3):
    rng = np.random.default_rng(seed)
    e_calc = np.linspace(1.5e-3, 19e-3, 2 * n_etrue)
    n_paths = n_cosz * n_sub
    seg_start, seg_layer, seg_length = [0], [], []
    for p in range(n_paths):
        depth = int(rng.integers(1, n_layers + 1))                 # deepest layer reached
        layers = list(range(1, depth + 1)) + list(range(depth - 1, 0, -1))
        seg_layer += layers
        seg_length += list(rng.uniform(50.0, 3000.0, len(layers)))
        seg_start.append(len(seg_layer))
    bg_es_par = np.array([1, 0][:n_bg_es] + [1] * max(0, n_bg_es - 2), dtype=np.int32)
    bg_cc_par = np.array(([1, 1, 0] * 3)[:n_bg_cc], dtype=np.int32)

    def resp(n):
        m = rng.uniform(0.0, 1.0, (n_etrue, n))
        return m / m.sum(axis=1, keepdims=True)

    n_norm = int(np.count_nonzero(bg_es_par)) + int(np.count_nonzero(bg_cc_par))
    truth = np.array([0.307, 0.0224, 7.5e-5, 5.2e6, 8.0e3, 1.0, 0.0, 0.0] + [1.0] * n_layers + [2.2e-6] * n_norm)
    return dict(
        n_cosz_sub=np.int32(n_sub), e_calc=e_calc, seg_start=np.array(seg_start, dtype=np.int32),
        seg_layer=np.array(seg_layer, dtype=np.int32), seg_length=np.array(seg_length),
        earth_lookup=np.sort(rng.uniform(1e-4, 8e-4, n_layers)), exposure_weights=rng.uniform(0.005, 0.03, n_cosz),
        avg_ne_8b=np.float64(89.0), avg_ne_hep=np.float64(70.0),
        solar_ne=np.linspace(100.0, 5.0, 37), solar_prod_8b=rng.uniform(0, 1, 37), solar_prod_hep=rng.uniform(0, 1, 37),
        unosc_es_nue_8b=rng.uniform(1e-4, 3e-4, n_etrue), unosc_es_nuother_8b=rng.uniform(2e-5, 6e-5, n_etrue),
        unosc_cc_8b=rng.uniform(1e-4, 5e-4, n_etrue), unosc_es_nue_hep=rng.uniform(1e-4, 3e-4, n_etrue),
        unosc_es_nuother_hep=rng.uniform(2e-5, 6e-5, n_etrue), unosc_cc_hep=rng.uniform(1e-4, 5e-4, n_etrue),
        log_e_norm=np.log(np.linspace(0.2, 2.0, n_etrue)),
        resp_es_nue=resp(n_es), resp_es_nuother=resp(n_es), resp_cc=resp(n_cc),
        eff_es_nue=rng.uniform(0.2, 0.9, n_es), eff_es_nuother=rng.uniform(0.2, 0.9, n_es), eff_cc=rng.uniform(0.2, 0.9, n_cc),
        raw_bg_es=rng.uniform(0.0, 2e-3, (n_bg_es, n_es)), bg_es_sides=np.array([0, 1, -1, 0][:n_bg_es], dtype=np.int32),
        bg_es_par=bg_es_par, raw_bg_cc=rng.uniform(0.0, 5e-3, (n_bg_cc, n_cc)), bg_cc_par=bg_cc_par,
        index_es=np.int32(idx_es), index_cc=np.int32(idx_cc), cc_xsec_norm_true=np.float64(1.0), truth=truth)


def draws(t, n, seed):
    rng = np.random.default_rng(seed)
    p = np.tile(np.asarray(t["truth"], dtype=float), (n, 1))
    nl = len(t["earth_lookup"])
    p[:, 0] = rng.uniform(0.2, 0.45, n); p[:, 1] = rng.uniform(0.01, 0.03, n); p[:, 2] = rng.uniform(3e-5, 2e-4, n)
    p[:, 3] *= rng.uniform(0.8, 1.2, n); p[:, 4] *= rng.uniform(0.5, 1.5, n)
    p[:, 5] = rng.uniform(0.8, 1.2, n); p[:, 6] = rng.normal(0, 0.1, n); p[:, 7] = rng.normal(0, 0.05, n)
    p[:, 8:8 + nl] = rng.uniform(0.6, 1.4, (n, nl))
    p[:, 8 + nl:] *= rng.uniform(0.9, 1.1, (n, p.shape[1] - 8 - nl))
    return p


SHAPES = [
    # n_etrue, n_cosz, n_sub, n_layers, n_es, n_cc, idx_es, idx_cc
    (7, 3, 1, 3, 5, 9, 2, 1),
    (18, 13, 2, 4, 17, 30, 4, 9),
    (40, 50, 1, 6, 30, 30, 7, 5),
    (33, 21, 3, 2, 8, 16, 1, 16),
    (4, 1, 1, 1, 1, 2, 1, 2),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fast", [True, False])
def test_other_binnings(shape, fast):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = synthetic_tables(17 + shape[0], *shape)
    data = O.asimov_data(t, fast=fast)
    p = draws(t, 4, seed=shape[1])
    with SolarOscEngine(t, data=data, fast=fast) as e:
        assert e.n_params == len(t["truth"])
        lnl = e.eval_llh(p)
        rates = e.eval_rates(p)
        big = e.eval_llh(np.tile(p, (50, 1)))                    # 200 points: the fused kernel where it fits
        a_es, a_cc = e.eval_asymmetry(np.tile(p, (50, 1)))
        osc = e.eval_oscprobs(p[:2])
    for i in range(len(p)):
        ref = O.propagate_samples(p[i], t, fast=fast)
        for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
            scale = np.maximum(np.abs(ref[k]), 1e-9 * np.max(np.abs(ref[k])))
            assert np.max(np.abs(rates[k][i] - ref[k]) / scale) < 1e-9, k
        assert abs(2 * lnl[i] - 2 * O.total_llh(p[i], t, data, fast=fast)) < 1e-6
    assert np.max(np.abs(big.reshape(50, -1) - lnl[None, :])) < 1e-8
    ro = O.compute_shared_osc_probs(p[0], t, fast=fast)
    assert np.max(np.abs(osc["nue_8B_night"][0] - ro["nue_8B_night"])) < 1e-11
    assert np.all(np.isfinite(a_es)) and np.all(np.isfinite(a_cc))


@pytest.mark.parametrize("incl,ang", [(1, 1), (2, 0), (2, 1)])
def test_other_binnings_of_the_branches(incl, ang):
    """The inclusive / semi-inclusive / angular fold on an odd binning (18 true-energy bins, 13 zenith bins with two
    sub-paths, 17 ES and 30 CC reco bins, 7 cos-scatter bins of which 3 are cut)."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    shape = (18, 13, 2, 4, 17, 30, 4, 9)
    t = synthetic_tables(99, *shape)
    rng = np.random.default_rng(7)
    n_etrue, n_es, n_cc, n_cos = 18, 17, 30, 7
    m = rng.uniform(0, 1, (n_etrue, n_es))
    t["resp_cc_incl"] = m / m.sum(axis=1, keepdims=True)
    t["eff_cc_incl"] = rng.uniform(0.3, 0.9, n_es)
    a = rng.uniform(0.1, 1.0, (n_cos, n_es))
    t["resp_es_angular"] = a / a.sum(axis=0, keepdims=True)
    t["angular_mask"] = np.array([0, 0, 0, 1, 1, 1, 1], dtype=float)
    t["resp_bg_angular"] = np.where(t["angular_mask"][:, None] > 0, 1.0 / 4, 0.0) * np.ones((n_cos, n_es))
    t["n_cos"] = np.int32(n_cos)
    t["resp_cc_misid_nue"] = 0.05 * rng.uniform(0, 1, (n_etrue, n_cc))
    t["resp_cc_misid_nuother"] = 0.05 * rng.uniform(0, 1, (n_etrue, n_cc))
    t["cc_f_above"], t["cc_f_below"] = np.float64(4 / 7), np.float64(3 / 7)
    t["inclusive"], t["angular"] = np.int32(incl), np.int32(ang)
    data = O.asimov_data(t)
    p = draws(t, 3, seed=incl + 2 * ang)
    with SolarOscEngine(t, data=data) as e:
        lnl = e.eval_llh(p)
        rates = e.eval_rates(p)
        big = e.eval_llh(np.tile(p, (60, 1)))                    # fused gen_llh_kernel
    for i in range(len(p)):
        ref = O.propagate_samples(p[i], t)
        for k in ("ES_day", "ES_night", "CC_day", "CC_night"):
            assert rates[k][i].shape == ref[k].shape, k
            scale = np.maximum(np.abs(ref[k]), 1e-9 * max(np.max(np.abs(ref[k])), 1e-300))
            assert np.max(np.abs(rates[k][i] - ref[k]) / np.where(scale > 0, scale, 1.0)) < 1e-9, k
        assert abs(2 * lnl[i] - 2 * O.total_llh(p[i], t, data)) < 1e-6
    assert np.max(np.abs(big.reshape(60, -1) - lnl[None, :])) < 1e-8


@pytest.mark.parametrize("shape", [SHAPES[1], SHAPES[3]])
def test_gradient_on_other_binnings(shape):
    """solarosc_eval_llh_grad on odd binnings: every column against a central difference of the value."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = synthetic_tables(31 + shape[0], *shape)
    data = O.asimov_data(t)
    p = draws(t, 3, seed=5)
    with SolarOscEngine(t, data=data) as e:
        lnl, g = e.eval_llh_grad(p)
        assert np.max(np.abs(lnl - e.eval_llh(p))) < 5e-7
        used = [c for c in range(p.shape[1])]
        for c in used:
            h = 1e-6 * max(abs(p[0, c]), 1e-3) if c != 2 else 1e-9
            pp, pm = p.copy(), p.copy()
            pp[:, c] += h; pm[:, c] -= h
            fd = (e.eval_llh(pp) - e.eval_llh(pm)) / (2 * h)
            scale = np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(g[:, c])) + 1e-12)
            assert np.max(np.abs(g[:, c] - fd) / scale) < 2e-3, (c, g[:, c], fd)
