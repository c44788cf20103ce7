"""total_llh sums the detectors of a run (likelihood_main.jl:30-35): shared oscillation / flux / cross-section / Earth
columns, background norms namespaced per detector (core.jl:6-15).  Two-detector handles against per-detector oracle
calls; edge batches (empty, single point)."""
import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(tables, asimov):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    with SolarOscEngine(tables, data=asimov) as e:
        yield e


def _second_detector(tables):
    """A second detector on the same shared tables: other exposures, thresholds and one fixed background component."""
    t = dict(tables)
    for k in ("unosc_es_nue_8b", "unosc_es_nuother_8b", "unosc_es_nue_hep", "unosc_es_nuother_hep"):
        t[k] = 2.0 * np.asarray(tables[k])                       # ES_exposure 40 instead of 20
    for k in ("unosc_cc_8b", "unosc_cc_hep"):
        t[k] = 4.0 * np.asarray(tables[k])                       # CC_exposure 40 instead of 10
    t["index_es"], t["index_cc"] = np.int32(9), np.int32(4)
    t["bg_es_par"] = np.array([1, 0, 1], dtype=np.int32)         # component 2 is not scaled by a nuisance
    return t


def _split(p, t0, t1):
    """Per-detector 19- and 18-column vectors of a two-detector parameter row."""
    nl = len(t0["earth_lookup"])
    k = 8 + nl
    n0 = int(np.count_nonzero(t0["bg_es_par"])) + int(np.count_nonzero(t0["bg_cc_par"]))
    return np.concatenate([p[:k], p[k:k + n0]]), np.concatenate([p[:k], p[k + n0:]])


def _draw2(tables, draw, t1, n, seed):
    base = draw(tables, n, seed=seed)
    n1 = int(np.count_nonzero(t1["bg_es_par"])) + int(np.count_nonzero(t1["bg_cc_par"]))
    extra = 2.2e-6 * np.random.default_rng(seed).normal(1.0, 0.02, (n, n1))
    return np.ascontiguousarray(np.hstack([base, extra]))


def test_two_detectors_sum(tables, asimov, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t1 = _second_detector(tables)
    d1 = O.asimov_data({**t1, "truth": np.concatenate([tables["truth"][:14], tables["truth"][14:16], tables["truth"][17:19]])})
    p = _draw2(tables, draw, t1, 5, seed=11)
    with SolarOscEngine(tables, detectors=[tables, t1], data=[asimov, d1], det_names=["HD", "VD"]) as e:
        assert e.n_params == 19 + 4 and e.param_names[-4:] == ["VD_ES_bg_norm_1", "VD_ES_bg_norm_2", "VD_CC_bg_norm_1", "VD_CC_bg_norm_2"]
        got = e.eval_llh(p)
        big = e.eval_llh(np.tile(p, (80, 1)))                   # enough points for the fused kernel
        r1 = e.eval_rates(p, det=1)
        g_l, g_g = e.eval_llh_grad(p)
    for i in range(len(p)):
        p0, p1 = _split(p[i], tables, t1)
        ref = O.total_llh(p0, tables, asimov) + O.total_llh(p1, t1, d1)
        assert abs(2 * got[i] - 2 * ref) < 1e-6
        rr = O.propagate_samples(p1, t1)
        for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
            assert np.max(np.abs(r1[k][i] - rr[k]) / np.maximum(np.abs(rr[k]), 1e-9 * np.max(np.abs(rr[k])))) < 1e-9, k
    assert np.max(np.abs(big.reshape(80, -1) - got[None, :])) < 1e-9
    assert np.max(np.abs(g_l - got)) < 2.5e-7 and g_g.shape == (5, 23)       # other launch shape of the oscillation stage: rounding
    # the gradient with respect to the second detector's CC norm, against a central difference of the value
    h = 1e-9
    pp, pm = p.copy(), p.copy()
    pp[:, -1] += h; pm[:, -1] -= h
    with SolarOscEngine(tables, detectors=[tables, t1], data=[asimov, d1]) as e:
        fd = (e.eval_llh(pp) - e.eval_llh(pm)) / (2 * h)
    assert np.max(np.abs(g_g[:, -1] - fd) / np.maximum(np.abs(fd), 1.0)) < 1e-4


def test_exclusive_and_angular_detectors_in_one_handle(tables, asimov, branches, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    tv = branches(tables, 1, 1)
    dv = O.asimov_data(tv)
    p = np.ascontiguousarray(np.hstack([draw(tables, 4, seed=12), np.tile(tables["truth"][14:19], (4, 1))]))
    with SolarOscEngine(tables, detectors=[tables, tv], data=[asimov, dv]) as e:
        got = e.eval_llh(p)
        r0, r1 = e.eval_rates(p, det=0), e.eval_rates(p, det=1)
        # value + gradient of the mixed handle (gen_grad_kernel walks both detectors): the value, and grad . v against
        # 4th-order differences of the handle's own lnL along random directions over all 24 columns
        lnl, grad = e.eval_llh_grad(p)
        assert np.max(np.abs(2 * lnl - 2 * got)) < 1e-6 and grad.shape == p.shape
        v = np.random.default_rng(4).normal(size=p.shape) * (np.abs(p) * 2e-4 + 1e-12)
        v[:, 14:] *= 1e-2
        f = [e.eval_llh(p + m * v) for m in (-2, -1, 1, 2)]
        fd = (f[0] - 8 * f[1] + 8 * f[2] - f[3]) / 12
        pred = np.sum(grad * v, axis=1)
        assert np.max(np.abs(fd - pred) / np.maximum(np.abs(pred), 1e-2)) < 1e-4
    assert r0["ES_night"].shape == (4, 50, 30) and r1["ES_night"].shape == (4, 25, 30, 50)
    for i in range(len(p)):
        p0 = p[i, :19]
        p1 = np.concatenate([p[i, :14], p[i, 19:]])
        ref = O.total_llh(p0, tables, asimov) + O.total_llh(p1, tv, dv)
        assert abs(2 * got[i] - 2 * ref) < 1e-6


def test_empty_and_single_point_batches(engine, tables):
    p = np.asarray(tables["truth"], dtype=float)[None, :]
    assert engine.eval_llh(np.empty((0, engine.n_params))).shape == (0,)
    assert engine.eval_rates(np.empty((0, engine.n_params)))["ES_night"].shape[0] == 0
    one = engine.eval_llh(p)
    assert one.shape == (1,) and abs(one[0]) < 1e-6             # Asimov data at the truth: zero deviance
    a_es, a_cc = engine.eval_asymmetry(p)
    assert a_es.shape == (1,) and np.isfinite(a_cc[0])


def test_one_process_several_handles(tables, asimov, draw):
    """solarosc_eval_llh_multi: the batch dealt over handles (one per GPU where the box has several; two handles on the
    one GPU otherwise), whole groups per slice, identical bits to a single handle."""
    import torch
    from dune_solaroscfitter_b200.engine import SolarOscEngine, eval_llh_multi
    n_dev = max(1, torch.cuda.device_count())
    devs = list(range(n_dev)) if n_dev > 1 else [0, 0]
    p = draw(tables, 27 * 701, seed=21)                       # 701 groups of 27: the slices cannot be equal
    p[27 * 50, 9] = 2.5                                       # one out-of-bounds earth norm
    engines = [SolarOscEngine(tables, data=asimov, device=d) for d in devs]
    try:
        one, f1 = engines[0].eval_llh(p, return_flags=True)
        many, fm = eval_llh_multi(engines, p, group=1, return_flags=True)
        assert np.array_equal(one, many) and np.array_equal(f1, fm) and many[27 * 50] == -np.inf
        g1 = engines[0].eval_llh(p, group=27)
        gm = eval_llh_multi(engines, p, group=27)
        assert np.array_equal(g1, gm)
        assert eval_llh_multi(engines, p[:0]).shape == (0,)
    finally:
        for e in engines:
            e.close()
