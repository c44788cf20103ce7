"""Inclusive, semi-inclusive and angular-ES branches of propagateSamples (propagation_main.jl:98-153,
propagation_reco.jl:66-199) and llh_ES_angle (likelihood_builder.jl:53-78).

CPU part: structural identities of the oracle's restatement.  GPU part: the general fold (gen_kernel) against the
oracle through the C ABI, every branch combination."""
import numpy as np
import pytest

from oracle import oracle_np as O


def _truth(t):
    return np.asarray(t["truth"], dtype=float)


def test_angular_rates_sum_back_to_the_energy_spectrum(tables, branches):
    """ES.angular and BG.angular columns sum to one over the cos bins, so without a cut the angular arrays summed
    over cos theta are the non-angular rates (propagation_reco.jl:86-96 vs :10-29)."""
    t = branches(tables, 0, 1)
    t["angular_mask"] = np.ones_like(t["angular_mask"])
    t["resp_bg_angular"] = np.full_like(t["resp_bg_angular"], 1.0 / int(t["n_cos"]))
    a = O.propagate_samples(_truth(t), t)
    b = O.propagate_samples(_truth(tables), tables)
    assert a["ES_day"].shape == (25, 30) and a["ES_night"].shape == (25, 30, 50)
    assert np.allclose(a["ES_day"].sum(axis=0), b["ES_day"], rtol=1e-12)
    assert np.allclose(a["ES_night"].sum(axis=0).T, b["ES_night"], rtol=1e-12)
    assert np.array_equal(a["CC_night"], b["CC_night"])


def test_angular_cut_zeroes_the_backward_bins(tables, branches):
    t = branches(tables, 0, 1)
    r = O.propagate_samples(_truth(t), t)
    back = t["angular_mask"] == 0
    assert back.sum() == 12 and np.all(r["ES_day"][back] == 0) and np.all(r["ES_night"][back] == 0)
    assert np.all(r["ES_day"][~back][:, 4:] > 0)          # (the efficiency is zero in the lowest reco bins)


def test_inclusive_moves_cc_into_the_es_bins(tables, branches):
    t = branches(tables, 1, 0)
    r = O.propagate_samples(_truth(t), t)
    assert np.all(r["CC_day"] == 0) and np.all(r["CC_night"] == 0)
    t0 = dict(t); t0["inclusive"] = np.int32(0)
    e = O.propagate_samples(_truth(t0), t0)
    extra = r["ES_day"] - e["ES_day"]
    assert np.all(extra >= 0) and extra.sum() > 0
    assert np.allclose(r["CC_incl_spectrum"], extra + (r["ES_night"] - e["ES_night"]).sum(axis=0), rtol=1e-9)
    # use_CC is false in the inclusive analysis (likelihood_main.jl:12): the CC data never enter
    d = {k: r[k] for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
    assert O.total_llh(_truth(t), t, d) == O.total_llh(_truth(t), t, {**d, "CC_day": d["CC_day"] + 5.0})


def test_semi_inclusive_splits_cc_between_the_hemispheres(tables, branches):
    t = branches(tables, 2, 0)
    r = O.propagate_samples(_truth(t), t)
    t0 = dict(t); t0["inclusive"] = np.int32(0)
    e = O.propagate_samples(_truth(t0), t0)
    t1 = dict(t); t1["inclusive"] = np.int32(1)
    i = O.propagate_samples(_truth(t1), t1)
    fa, fb = float(t["cc_f_above"]), float(t["cc_f_below"])
    assert abs(fa + fb - 1) < 1e-15
    assert np.allclose(r["ES_day"] - e["ES_day"], fa * (i["ES_day"] - e["ES_day"]), rtol=1e-9, atol=1e-12)
    misid = r["CC_day"] - fb * e["CC_day"]
    assert np.all(misid >= -1e-12) and misid.sum() > 0      # backward-scattered ES read as CC


# ---- GPU parity ------------------------------------------------------------------------------------------------------
BRANCHES = [(1, 0), (2, 0), (0, 1), (1, 1), (2, 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("incl,ang", BRANCHES)
def test_general_fold_matches_the_oracle(tables, branches, draw, incl, ang):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = branches(tables, incl, ang)
    data = O.asimov_data(t)
    p = draw(tables, 6, seed=300 + 10 * incl + ang)
    with SolarOscEngine(t, data=data) as e:
        got = e.eval_rates(p)
        lnl = e.eval_llh(p)
        many = e.eval_llh(np.tile(p, (60, 1)))          # several chunks' worth of CTAs
        spec = e.eval_cc_incl_spectrum(p)
    for i in range(len(p)):
        ref = O.propagate_samples(p[i], t)
        if incl == 0:                                   # CC_incl_spectrum: `nothing` in the exclusive analysis
            assert spec is None and ref["CC_incl_spectrum"] is None
        else:                                           # propagation_main.jl:108-146
            assert np.max(np.abs(spec[i] - ref["CC_incl_spectrum"]) / np.maximum(np.abs(ref["CC_incl_spectrum"]), 1e-300)) < 1e-9
        for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
            a, b = got[k][i], ref[k]
            assert a.shape == b.shape, (k, a.shape, b.shape)
            scale = np.maximum(np.abs(b), 1e-9 * np.max(np.abs(b)))
            assert np.max(np.abs(a - b) / np.where(scale > 0, scale, 1.0)) < 1e-9, k     # BASELINE.json: rates <= 1e-9 relative
        assert abs(2 * lnl[i] - 2 * O.total_llh(p[i], t, data)) < 1e-6                  # and -2lnL <= 1e-6 absolute
    assert np.array_equal(many.reshape(60, -1), np.tile(lnl, (60, 1)))


@pytest.mark.gpu
def test_general_fold_julia_layout_and_unsupported_calls(tables, branches, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = branches(tables, 1, 1)
    data = O.asimov_data(t)
    p = draw(tables, 3, seed=77)
    with SolarOscEngine(t, data=data) as e:
        a = e.eval_rates(p)
        la = e.eval_llh(p)
    # the same tables and data as Julia lays them out (column-major: numpy views with reversed shapes)
    tj = dict(t)
    for k in ("resp_es_nue", "resp_es_nuother", "resp_cc", "resp_cc_incl", "resp_es_angular", "resp_bg_angular"):
        tj[k] = np.ascontiguousarray(np.asarray(t[k]).T)
    dj = {k: np.ascontiguousarray(np.transpose(v)) for k, v in data.items()}
    with SolarOscEngine(tj, data=dj, col_major=True) as e:
        b = e.eval_rates(p)
        lb = e.eval_llh(p)
    assert np.array_equal(la, lb)
    assert np.array_equal(b["ES_day"], np.transpose(a["ES_day"], (0, 2, 1)))
    assert np.array_equal(b["ES_night"], np.transpose(a["ES_night"], (0, 3, 2, 1)))
    assert np.array_equal(b["CC_night"], np.transpose(a["CC_night"], (0, 2, 1)))


def _perbin_ref(r, m, idx, axis):
    """per-bin terms on the energy bins idx..n-1 (overflow bin excluded), zeros elsewhere
    (likelihood_builder.jl:22-51, 80-108, 146-174)."""
    out = np.zeros_like(r)
    n = r.shape[axis]
    sl = [slice(None)] * r.ndim
    sl[axis] = slice(idx - 1, n - 1)
    sl = tuple(sl)
    out[sl] = O.perbin_poisson_log_likelihood(r[sl], m[sl])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("incl,ang", [(1, 0), (2, 0), (0, 1), (1, 1), (2, 1)])
def test_perbin_and_asymmetry_of_the_branches(tables, branches, draw, incl, ang):
    """make_perbin_likelihood with llh_ES_angle_perbin / llh_ES_poisson_perbin, and the day-night asymmetries of
    derive_variables_from_chain.jl:236-258 (all cos-scatter bins summed; no CC asymmetry in the inclusive analysis)."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = branches(tables, incl, ang)
    data = O.asimov_data(t)
    p = draw(tables, 5, seed=400 + 10 * incl + ang)
    p[4, 9] = 2.2                                              # one out-of-bounds earth norm: per-bin terms are -Inf
    with SolarOscEngine(t, data=data) as e:
        pb = e.eval_perbin(p)
        a_es, a_cc = e.eval_asymmetry(p)
    idx_es, idx_cc = int(t["index_es"]), int(t["index_cc"])
    for i in range(4):
        r = O.propagate_samples(p[i], t)
        ref = dict(ES_day=_perbin_ref(r["ES_day"], data["ES_day"], idx_es, 1 if ang else 0),
                   ES_night=_perbin_ref(r["ES_night"], data["ES_night"], idx_es, 1),
                   CC_day=_perbin_ref(r["CC_day"], data["CC_day"], idx_cc, 0),
                   CC_night=_perbin_ref(r["CC_night"], data["CC_night"], idx_cc, 1))
        if incl == 1:                                          # use_CC is false: the CC arrays stay zero
            ref["CC_day"][:], ref["CC_night"][:] = 0.0, 0.0
        for k, b in ref.items():
            a = pb[k][i]
            assert a.shape == b.shape, k
            assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)) < 1e-9, k
        bg_es = r["BG_ES_tot"][idx_es - 1:].sum()
        d = (r["ES_day"][:, idx_es - 1:] if ang else r["ES_day"][idx_es - 1:]).sum() - 0.5 * bg_es
        n = r["ES_night"][:, idx_es - 1:].sum() - 0.5 * bg_es
        assert abs(a_es[i] - 2 * (d - n) / (d + n)) < 1e-9
        if incl == 1:
            assert a_cc[i] == 0.0
        else:
            bg_cc = r["BG_CC_tot"][idx_cc - 1:].sum()
            d = r["CC_day"][idx_cc - 1:].sum() - 0.5 * bg_cc
            n = r["CC_night"][:, idx_cc - 1:].sum() - 0.5 * bg_cc
            assert abs(a_cc[i] - 2 * (d - n) / (d + n)) < 1e-9
    assert np.all(pb["ES_night"][4] == -np.inf) and np.all(pb["ES_day"][4] == -np.inf)


@pytest.mark.gpu
@pytest.mark.parametrize("incl", [1, 2])
def test_derived_sums_of_the_non_angular_branches(tables, branches, draw, incl):
    """solarosc_eval_derived on an inclusive / semi-inclusive detector: the rate moments are the branch's rates."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = branches(tables, incl, 0)
    data = O.asimov_data(t)
    p = draw(tables, 40, seed=500 + incl)
    w = np.random.default_rng(incl).integers(1, 4, 40).astype(float)
    with SolarOscEngine(t, data=data) as e:
        got = e.eval_derived(p, w)
        rates = e.eval_rates(p)
        pb = e.eval_perbin(p)
    k = int(t["index_es"]) - 1
    m1 = np.einsum("i,ijk->jk", w, rates["ES_night"])
    assert np.max(np.abs(got["mom_ES"][0][1:, k:] - m1[:, k:]) / m1[:, k:]) < 1e-12
    assert np.all(got["mom_ES"][0][:, :k] == 0)
    s1 = np.einsum("i,ijk->jk", w, pb["CC_night"])
    assert np.max(np.abs(got["mom_CC"][2][1:] - s1)) <= 1e-9 * max(1.0, np.max(np.abs(s1)))
    if incl == 1:
        assert np.all(got["mom_CC"][:2] == 0)                   # the CC arrays are zero in the inclusive analysis


@pytest.mark.gpu
@pytest.mark.parametrize("incl,col_major", [(0, False), (1, False), (2, True)])
def test_derived_sums_of_the_angular_arrays(tables, branches, draw, incl, col_major):
    """solarosc_eval_derived with the angular ES reconstruction (derive_variables_from_chain.jl:239-258, 288-296, 315-318):
    the device block sums over the (n_cos, n_ereco[, n_cosz]) arrays against the replay loop written out in numpy on the
    engine's own per-sample rates and per-bin terms (which test_general_fold_matches_the_oracle pins to the oracle)."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle.derived_np import replay_block_sums_angular
    t = branches(tables, incl, 1)
    data = O.asimov_data(t)
    n = 600                                                # three device chunks of the reduction (max_batch = 256 below)
    p = draw(tables, n, seed=900 + incl)
    p[7, 8] = 2.5                                          # one sample outside the earth-norm bounds: -Inf per-bin terms
    w = np.random.default_rng(incl).integers(1, 4, n).astype(float)
    tj, dj = t, data
    if col_major:
        tj = dict(t)
        for k in ("resp_es_nue", "resp_es_nuother", "resp_cc", "resp_cc_incl", "resp_es_angular", "resp_bg_angular",
                  "resp_cc_misid_nue", "resp_cc_misid_nuother"):
            if k in t:
                tj[k] = np.ascontiguousarray(np.asarray(t[k]).T)
        dj = {k: np.ascontiguousarray(np.transpose(v)) for k, v in data.items()}
    with SolarOscEngine(t, data=data) as e0:               # row-major twin: per-sample arrays for the replay loop
        rates, pb = e0.eval_rates(p), e0.eval_perbin(p)
        lref = {k: v[0] for k, v in e0.eval_perbin(np.asarray(t["truth"], dtype=float)[None, :]).items()}
    with SolarOscEngine(tj, data=dj, col_major=col_major, max_batch=256) as e:
        lr = lref if not col_major else {k: np.ascontiguousarray(np.transpose(v)) for k, v in lref.items()}
        got = e.eval_derived(p, w, lr)
        a_es, _ = e.eval_asymmetry(p[:64])
    ok = np.ones(n, dtype=bool); ok[7] = False
    sub = lambda d: {k: v[ok] for k, v in d.items()}     # noqa: E731  (the -Inf sample poisons the sums by design: checked apart)
    ref = replay_block_sums_angular(rates, pb, lref, p, w, int(t["index_es"]))
    assert np.array_equal(got["ES_day"], ref["ES_day"]) or np.max(np.abs(got["ES_day"] - ref["ES_day"])) <= 1e-12 * np.max(ref["ES_day"])
    assert np.all(np.isneginf(got["ltot"][7, :2])) and np.allclose(got["ltot"][ok, :2], ref["ltot_ES"][ok], rtol=1e-11, atol=1e-9)
    for q in (0, 1):                                       # rate moments: every sample counts, zero below the threshold
        for key in ("mom_ES_day", "mom_ES_night"):
            r = ref[key][q]
            assert np.max(np.abs(got[key][q] - r)) <= 1e-12 * np.max(np.abs(r)), (key, q)
    k = int(t["index_es"]) - 1
    assert np.all(got["mom_ES_day"][0][:, :k] == 0) and np.all(got["mom_ES_night"][1][:, :k, :] == 0)
    assert np.all(np.isinf(got["mom_ES_day"][2]) | np.isnan(got["mom_ES_day"][2]))     # the out-of-bounds sample, as in the reference
    with SolarOscEngine(t, data=data) as e1:               # the same block without that sample: finite centred sums
        g1 = e1.eval_derived(p[ok], w[ok], lref)
    r1 = replay_block_sums_angular(sub(rates), sub(pb), lref, p[ok], w[ok], int(t["index_es"]))
    for q in range(2, 6):
        for key in ("mom_ES_day", "mom_ES_night"):
            r = r1[key][q]
            assert np.max(np.abs(g1[key][q] - r)) <= 1e-11 * max(1.0, np.max(np.abs(r))), (key, q)
    # asymmetry of the angular arrays (derive...:239-246): sums over every cos-scatter bin from the threshold on
    bg = rates["BG_ES_tot"][:64, k:].sum(axis=1)
    dd = rates["ES_day"][:64][:, :, k:].sum(axis=(1, 2)) - 0.5 * bg
    nn = rates["ES_night"][:64][:, :, k:, :].sum(axis=(1, 2, 3)) - 0.5 * bg
    assert np.max(np.abs(a_es - 2 * (dd - nn) / (dd + nn))) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("incl,ang", BRANCHES)
def test_gradient_of_the_branches(tables, branches, draw, incl, ang):
    """solarosc_eval_llh_grad on the inclusive / semi-inclusive / angular branches (gen_grad_kernel + adjoint_kernel):
    the value is eval_llh's, every column agrees with 4th-order central differences of the engine's own lnL (which
    test_general_fold_matches_the_oracle pins to the oracle) along the coordinate axes and along random directions."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    t = branches(tables, incl, ang)
    data = O.asimov_data(t)
    p = draw(tables, 12, seed=700 + 10 * incl + ang)
    p[0] = t["truth"]; p[0, 0] += 0.01; p[0, 3] *= 1.02
    p[11, 9] = 2.3                                         # out of the earth-norm bounds: -Inf, zero gradient row
    with SolarOscEngine(t, data=data) as e:
        lnl, grad = e.eval_llh_grad(p)
        ref = e.eval_llh(p)
        assert np.isneginf(lnl[11]) and np.all(grad[11] == 0)
        ok = np.arange(11)
        assert np.max(np.abs(2 * lnl[ok] - 2 * ref[ok])) < 1e-6
        assert np.all(np.isfinite(grad[ok]))
        P = p.shape[1]
        for k in range(P):
            if k in (15, 16):                              # ES background norms 2 and 3 do not enter (every ES component uses norm_1)
                assert np.all(grad[ok, k] == 0)
                continue
            if k == 2:
                h = 1e-4 * np.abs(p[ok, k])                 # dm2_21: the phases make lnL oscillate in it
            elif k in (0, 1) or 8 <= k < 14:
                h = np.full(len(ok), 1e-3)                  # mixing angles, earth norms
            elif k >= 14:
                h = 1e-3 * np.abs(p[ok, k])                 # background norms (~1e-5): lnL is strongly curved in them
            else:
                h = 1e-3 * np.maximum(np.abs(p[ok, k]), 1e-2)   # fluxes, cross-section norm / tilt / curvature
            f = []
            for m in (-2, -1, 1, 2):
                q = p[ok].copy(); q[:, k] += m * h
                f.append(e.eval_llh(q))
            fd = (f[0] - 8 * f[1] + 8 * f[2] - f[3]) / (12 * h)
            tol = (2e-4 if k == 2 else 2e-5) * np.abs(grad[ok, k]) + 2e-6 / h
            assert np.all(np.abs(fd - grad[ok, k]) <= tol), (k, fd, grad[ok, k])
        rng = np.random.default_rng(incl + 3 * ang)
        v = rng.normal(size=p[ok].shape) * (np.abs(p[ok]) * 2e-4 + 1e-12)
        v[:, 14:] *= 1e-2
        f = [e.eval_llh(p[ok] + m * v) for m in (-2, -1, 1, 2)]
        fd = (f[0] - 8 * f[1] + 8 * f[2] - f[3]) / 12
        pred = np.sum(grad[ok] * v, axis=1)
        # (the differences carry the rounding noise of up to 38 k Poisson terms per point with the angular reconstruction)
        assert np.max(np.abs(fd - pred) / np.maximum(np.abs(pred), 1e-2)) < (1e-4 if ang else 5e-6)
