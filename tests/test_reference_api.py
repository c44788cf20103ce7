"""Host-side mirror of the reference interface: parameter packing and wiring (no GPU needed for these)."""
import numpy as np
import pytest

from dune_solaroscfitter_b200 import reference_api as R


def test_params_by_prefix_sorts_names_as_strings():
    p = {"earth_norm_2": 2.0, "earth_norm_10": 10.0, "earth_norm_1": 1.0, "other": 0}
    assert R.params_by_prefix(p, "earth_norm_") == (1.0, 10.0, 2.0)       # core.jl:8-11 quirk: "10" < "2"
    assert R.params_by_prefix(p, "nothing_") == ()


def test_check_earth_norm_bounds():
    # reference tests/test_likelihood_correctness.jl:53-67
    assert R.check_earth_norm_bounds({"earth_norm_1": 0.5, "earth_norm_2": 1.9, "other": 3.0})
    assert not R.check_earth_norm_bounds({"earth_norm_1": -0.01})
    assert not R.check_earth_norm_bounds({"earth_norm_1": 2.01})
    assert R.check_earth_norm_bounds({"theta": 1.0})


class _FakeSession(R.SolarOscSession):
    def __init__(self):          # packing logic only, no engine
        self.det_names, self.earth_uncertainty, self.cc_xsec_norm_true = ["HD"], True, 1.0
        self.n_layers, self.n_norms = 6, [(3, 2)]


def _truth():
    p = dict(sin2_th12=0.307, sin2_th13=0.0224, dm2_21=6e-5, integrated_8B_flux=5.25e6, integrated_HEP_flux=8250.0,
             ES_asymmetry=0, CC_asymmetry=0)
    p.update({f"earth_norm_{i}": 1.0 + 0.01 * i for i in range(1, 7)})
    p.update({f"HD_ES_bg_norm_{i}": 2.2e-6 * i for i in range(1, 4)})
    p.update({f"HD_CC_bg_norm_{i}": 3.3e-6 * i for i in range(1, 3)})
    return p


def test_pack_order_defaults_and_ignored_fields():
    s = _FakeSession()
    row = s.pack(_truth())
    assert row.shape == (19,)
    assert row[:5].tolist() == [0.307, 0.0224, 6e-5, 5.25e6, 8250.0]
    assert row[5:8].tolist() == [1.0, 0.0, 0.0]                 # cc_xsec_* defaults (propagation_osc.jl:185-187)
    assert np.allclose(row[8:14], [1.01, 1.02, 1.03, 1.04, 1.05, 1.06])
    assert np.allclose(row[14:17], [2.2e-6, 4.4e-6, 6.6e-6]) and np.allclose(row[17:], [3.3e-6, 6.6e-6])
    p = _truth(); p.update(cc_xsec_norm=1.1, cc_xsec_tilt=0.05, cc_xsec_curv=-0.02)
    assert s.pack(p)[5:8].tolist() == [1.1, 0.05, -0.02]
    p = _truth(); del p["earth_norm_3"]
    with pytest.raises(KeyError):
        s.pack(p)
    assert s.pack_many([_truth(), _truth()]).shape == (2, 19)


@pytest.mark.gpu
def test_session_matches_oracle_through_the_reference_surface(tables, asimov):
    from oracle import oracle_np as O
    s = R.SolarOscSession(tables, det_names=["HD"], data=[asimov])
    p = _truth()
    row = s.pack(p)
    es_d, cc_d, es_n, cc_n, bg_es, bg_cc, incl = s.propagateSamples(params=p, det_flags={"det_name": "HD"})
    ref = O.propagate_samples(row, tables)
    assert incl is None
    for got, k in ((es_d, "ES_day"), (cc_d, "CC_day"), (es_n, "ES_night"), (cc_n, "CC_night"), (bg_es, "BG_ES_tot"), (bg_cc, "BG_CC_tot")):
        assert np.max(np.abs(got - ref[k]) / np.maximum(np.abs(ref[k]), 1e-300)) < 1e-9
    li = s.likelihood_inputs(0, asimov)
    r = R.expected_rates(li, p)
    assert np.array_equal(r["CC_night"], cc_n) and li.index_ES == int(tables["index_es"])
    llh = R.make_likelihood(s, li)
    assert abs(2 * llh(p) - 2 * O.total_llh(row, tables, asimov)) < 1e-6
    bad = dict(p); bad["earth_norm_2"] = 2.5
    assert llh(bad) == -np.inf
    osc = s.compute_shared_osc_probs(p)
    on = O.compute_shared_osc_probs(row, tables)
    assert osc["nue_8B_night"].shape == (50, 40) and np.max(np.abs(osc["nue_8B_night"] - on["nue_8B_night"])) < 5e-12
    assert np.allclose(osc["nuother_hep_day"], 1 - on["nue_hep_day"])
    s.close()
