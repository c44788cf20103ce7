"""The parity-pinning kit (julia/dump_fixtures.jl -> tools/check_against_julia.py) exercised mechanically: a dump in
the Julia script's format (raw column-major arrays + text manifest) is written HERE from the oracle's own outputs, read
back and checked.  This pins the file format, the table mapping and the checker's logic -- not parity: a dump written
by the reference itself (Julia) is what pins parity, and tests/test_julia_fixture.py consumes one when it is committed."""
import os
import subprocess
import sys

import numpy as np

from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Writer:
    def __init__(self, prefix):
        self.bin, self.txt = open(prefix + ".bin", "wb"), open(prefix + ".txt", "w")

    def put(self, name, a, dtype="Float64"):
        np_dt = {"Float64": "<f8", "Int32": "<i4", "UInt8": "u1"}[dtype]
        a = np.atleast_1d(np.asarray(a)).astype(np_dt)
        self.txt.write(" ".join([name, dtype, str(self.bin.tell()), str(a.ndim)] + [str(s) for s in a.shape]) + "\n")
        self.bin.write(np.asfortranarray(a).tobytes(order="F"))          # Julia memory order

    def close(self):
        self.bin.close(); self.txt.close()


def write_dump(prefix, t, data, params, perturb=0.0):
    w = Writer(prefix)
    for k in ("e_calc", "seg_length", "earth_lookup", "exposure_weights", "solar_ne", "solar_prod_8b", "solar_prod_hep"):
        w.put(k, t[k])
    w.put("seg_start", t["seg_start"], "Int32"); w.put("seg_layer", t["seg_layer"], "Int32")
    for k, v in (("n_cosz_sub", t["n_cosz_sub"]), ("avg_ne_8b", t["avg_ne_8b"]), ("avg_ne_hep", t["avg_ne_hep"]), ("fast", 1.0),
                 ("earth_uncertainty", 1.0), ("cc_xsec_norm_true", 1.0), ("n_det", 1)):
        w.put(k, float(v))
    w.put("det0/name", np.frombuffer(b"HD", dtype=np.uint8), "UInt8")
    for k, v in (("es_mode", 1.0), ("cc_mode", 1.0), ("index_es", t["index_es"]), ("index_cc", t["index_cc"]), ("inclusive", 0), ("angular", 0)):
        w.put("det0/" + k, float(v))
    for k in ("unosc_es_nue_8b", "unosc_es_nuother_8b", "unosc_cc_8b", "unosc_es_nue_hep", "unosc_es_nuother_hep", "unosc_cc_hep",
              "log_e_norm", "resp_es_nue", "resp_es_nuother", "resp_cc", "eff_es_nue", "eff_es_nuother", "eff_cc"):
        w.put("det0/" + k, t[k])
    w.put("det0/raw_bg_es", np.asarray(t["raw_bg_es"]).T); w.put("det0/raw_bg_cc", np.asarray(t["raw_bg_cc"]).T)   # (n_Ereco, components) in Julia
    for k in ("bg_es_sides", "bg_es_par", "bg_cc_par"):
        w.put("det0/" + k, t[k], "Int32")
    for k, src in (("obs_es_day", "ES_day"), ("obs_es_night", "ES_night"), ("obs_cc_day", "CC_day"), ("obs_cc_night", "CC_night")):
        w.put("det0/" + k, data[src])
    w.put("n_points", float(len(params)))
    w.put("params", params.T)
    w.put("total_llh", [O.total_llh(p, t, data) + perturb for p in params])
    for i, p in enumerate(params[:2]):
        osc = O.compute_shared_osc_probs(p, t)
        for k in ("nue_8B_day", "nue_8B_night", "nue_hep_day", "nue_hep_night"):
            w.put(f"point{i}/{k}", osc[k])
        r = O.propagate_samples(p, t)
        for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
            w.put(f"point{i}/det0/{k}", r[k])
    mp = O.get_mixing_parameters(O.view(params[0], t))
    w.put("point0/p1e", O.osc_prob_earth(t["e_calc"], mp, params[0][8:14] * t["earth_lookup"], t["seg_start"], t["seg_layer"], t["seg_length"]))
    w.close()


def run_checker(prefix):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_against_julia.py"), prefix, "--no-gpu"],
                          capture_output=True, text=True, cwd=ROOT)


def test_dump_format_round_trip_and_checker(tmp_path, tables, asimov, draw):
    p = draw(tables, 4, seed=12)
    p[0] = tables["truth"]
    p[3, 9] = 2.5                                   # out of bounds: -Inf
    prefix = str(tmp_path / "fixture")
    write_dump(prefix, tables, asimov, p)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_against_julia as cj
    d = cj.load_dump(prefix)
    assert d["params"].shape == (p.shape[1], 4) and np.array_equal(d["params"].T, p)
    assert d["point0/nue_8B_night"].shape == (50, 40) and d["det0/resp_cc"].shape == (40, 30)
    shared, dets, data = cj.tables_from_dump(d)
    assert np.array_equal(dets[0]["raw_bg_cc"], tables["raw_bg_cc"]) and np.array_equal(data[0]["CC_night"], asimov["CC_night"])
    r = run_checker(prefix)
    assert r.returncode == 0 and "PARITY PINNED" in r.stdout, r.stdout + r.stderr


def test_checker_reports_a_deviation(tmp_path, tables, asimov, draw):
    p = draw(tables, 2, seed=13)
    prefix = str(tmp_path / "bad")
    write_dump(prefix, tables, asimov, p, perturb=1e-5)     # -2lnL off by 2e-5: outside the 1e-6 tolerance
    r = run_checker(prefix)
    assert r.returncode == 1 and "FAILED" in r.stdout
