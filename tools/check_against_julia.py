#!/usr/bin/env python
"""Parity-pinning kit, Python side: check the oracle (always) and the engine (when a B200 is present) against a dump of
the REFERENCE's own results written by julia/dump_fixtures.jl inside a SolarOscFitter session.

    python tools/check_against_julia.py <prefix> [--no-gpu] [--save tests/golden/julia_fixture.npz]

Tolerances (BASELINE.json): per-bin rates <= 1e-9 relative, -2lnL <= 1e-6 absolute, oscillation probabilities and P1e
<= 5e-12 absolute, per-bin terms <= 1e-9 relative.  Everything is evaluated on the DUMPED tables, so set-up quadratures
(QuadGK) play no role.  --save stores tables + expected results as one .npz for tests/ (tests/test_julia_fixture.py picks
up tests/golden/julia_fixture*.npz when present).  Single-detector dumps are checked array by array; for several
detectors the per-detector rates and the summed lnL.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_DTYPES = {"Float64": "<f8", "Int32": "<i4", "Int64": "<i8", "UInt8": "u1"}


def load_dump(prefix: str) -> dict:
    """name -> array (Julia column-major arrays come back with their Julia shape)."""
    raw = np.fromfile(prefix + ".bin", dtype=np.uint8)
    out = {}
    for line in open(prefix + ".txt"):
        f = line.split()
        name, dt, off, nd = f[0], np.dtype(_DTYPES[f[1]]), int(f[2]), int(f[3])
        shape = tuple(int(x) for x in f[4:4 + nd])
        n = int(np.prod(shape)) if shape else 1
        a = raw[off:off + n * dt.itemsize].view(dt)
        out[name] = a.reshape(shape, order="F") if nd > 1 else a.copy()
    return out


def scalar(d, k):
    return float(d[k][0])


def tables_from_dump(d: dict):
    """(shared tables, [detector tables], [data]) in the layout of tests/golden/tables_escc.npz (row-major numpy)."""
    shared = {k: d[k] for k in ("e_calc", "seg_start", "seg_layer", "seg_length", "earth_lookup", "exposure_weights",
                                "solar_ne", "solar_prod_8b", "solar_prod_hep")}
    shared["n_cosz_sub"] = np.int32(scalar(d, "n_cosz_sub"))
    shared["avg_ne_8b"], shared["avg_ne_hep"] = np.float64(scalar(d, "avg_ne_8b")), np.float64(scalar(d, "avg_ne_hep"))
    shared["cc_xsec_norm_true"] = np.float64(scalar(d, "cc_xsec_norm_true"))
    dets, data = [], []
    for i in range(int(scalar(d, "n_det"))):
        pre = f"det{i}/"
        t = {k: np.ascontiguousarray(d[pre + k]) for k in ("unosc_es_nue_8b", "unosc_es_nuother_8b", "unosc_cc_8b", "unosc_es_nue_hep",
                                                            "unosc_es_nuother_hep", "unosc_cc_hep", "log_e_norm", "resp_es_nue",
                                                            "resp_es_nuother", "resp_cc", "eff_es_nue", "eff_es_nuother", "eff_cc")}
        t["raw_bg_es"] = np.ascontiguousarray(d[pre + "raw_bg_es"].T)          # (components, n_Ereco)
        t["raw_bg_cc"] = np.ascontiguousarray(d[pre + "raw_bg_cc"].T)
        for k in ("bg_es_sides", "bg_es_par", "bg_cc_par"):
            t[k] = d[pre + k].astype(np.int32)
        t["index_es"], t["index_cc"] = np.int32(scalar(d, pre + "index_es")), np.int32(scalar(d, pre + "index_cc"))
        t["es_mode"], t["cc_mode"] = bool(scalar(d, pre + "es_mode")), bool(scalar(d, pre + "cc_mode"))
        t["inclusive"], t["angular"] = np.int32(scalar(d, pre + "inclusive")), np.int32(scalar(d, pre + "angular"))
        for k in ("resp_cc_incl", "eff_cc_incl", "resp_cc_misid_nue", "resp_cc_misid_nuother", "resp_es_angular", "resp_bg_angular", "angular_mask"):
            if pre + k in d:
                t[k] = np.ascontiguousarray(d[pre + k])
        for k in ("cc_f_above", "cc_f_below", "n_cos"):
            if pre + k in d:
                t[k] = scalar(d, pre + k) if k != "n_cos" else np.int32(scalar(d, pre + k))
        t["name"] = bytes(d[pre + "name"]).decode()
        dets.append(t)
        data.append({"ES_day": d[pre + "obs_es_day"], "ES_night": np.ascontiguousarray(d[pre + "obs_es_night"]),
                     "CC_day": d[pre + "obs_cc_day"], "CC_night": np.ascontiguousarray(d[pre + "obs_cc_night"])})
    return shared, dets, data


def rel(a, b):
    b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), 1e-12 * np.max(np.abs(b)) + 1e-300)
    return float(np.max(np.abs(np.asarray(a) - b) / scale))


def check(name, err, tol, failures):
    ok = err <= tol
    print(f"  {'ok  ' if ok else 'FAIL'} {name:48s} {err:.3e}  (tolerance {tol:.0e})")
    if not ok:
        failures.append(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--save")
    a = ap.parse_args()
    d = load_dump(a.prefix)
    shared, dets, data = tables_from_dump(d)
    fast, earth_unc = bool(scalar(d, "fast")), bool(scalar(d, "earth_uncertainty"))
    params = np.ascontiguousarray(d["params"].T)                                  # (n_points, P)
    ref_lnl = d["total_llh"]
    n_arr = sum(1 for k in d if k.endswith("/nue_8B_day"))
    failures = []

    from oracle import oracle_np
    print(f"{len(params)} points, {len(dets)} detector(s), fast = {fast}, earth_uncertainty = {earth_unc}")
    print("oracle (numpy restatement) against the reference:")
    tabs = [dict(shared, **t) for t in dets]
    lnl = np.array([sum(oracle_np.total_llh(p, t, dat, fast, earth_unc, es_mode=t["es_mode"], cc_mode=t["cc_mode"])
                        for t, dat in zip(tabs, data)) for p in params])
    fin = np.isfinite(ref_lnl)
    check("-2lnL (finite points)", float(np.max(np.abs(2 * lnl[fin] - 2 * ref_lnl[fin]))), 1e-6, failures)
    check("-Inf on earth-norm bounds", float(np.any(np.isinf(ref_lnl) != np.isinf(lnl))), 0.0, failures)
    for i in range(n_arr):
        osc = oracle_np.compute_shared_osc_probs(params[i], tabs[0], fast, earth_unc)
        for k in ("nue_8B_day", "nue_8B_night", "nue_hep_day", "nue_hep_night"):
            check(f"point {i} {k}", float(np.max(np.abs(osc[k] - d[f'point{i}/{k}']))), 5e-12, failures)
        for j, t in enumerate(tabs):
            r = oracle_np.propagate_samples(params[i], t, fast, earth_unc, es_mode=t["es_mode"], cc_mode=t["cc_mode"])
            for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
                check(f"point {i} det {j} {k}", rel(r[k], d[f"point{i}/det{j}/{k}"]), 1e-9, failures)
    if "point0/p1e" in d:
        mp = oracle_np.get_mixing_parameters(oracle_np.view(params[0], tabs[0]))
        lookup = (params[0][8:8 + len(shared["earth_lookup"])] if earth_unc else 1.0) * shared["earth_lookup"]
        p1e = oracle_np.osc_prob_earth(shared["e_calc"], mp, lookup, shared["seg_start"], shared["seg_layer"], shared["seg_length"])
        check("point 0 P1e", float(np.max(np.abs(p1e - d["point0/p1e"]))), 5e-12, failures)

    import torch
    if not a.no_gpu and torch.cuda.is_available():
        from dune_solaroscfitter_b200.engine import SolarOscEngine
        print("engine (libsolarosc_b200.so) against the reference:")
        with SolarOscEngine(shared, dets, data, fast=fast, earth_uncertainty=earth_unc) as e:
            got, fl = e.eval_llh(params, return_flags=True)
            check("-2lnL (finite points)", float(np.max(np.abs(2 * got[fin] - 2 * ref_lnl[fin]))), 1e-6, failures)
            check("-Inf on earth-norm bounds", float(np.any(np.isinf(ref_lnl) != np.isinf(got))), 0.0, failures)
            osc = e.eval_oscprobs(params[:n_arr])
            for i in range(n_arr):
                for k in ("nue_8B_day", "nue_8B_night", "nue_hep_day", "nue_hep_night"):
                    check(f"point {i} {k}", float(np.max(np.abs(osc[k][i] - d[f'point{i}/{k}']))), 5e-12, failures)
            for j in range(len(dets)):
                r = e.eval_rates(params[:n_arr], det=j)
                for i in range(n_arr):
                    for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
                        check(f"point {i} det {j} {k}", rel(r[k][i], d[f"point{i}/det{j}/{k}"]), 1e-9, failures)
            if "point0/p1e" in d:
                check("point 0 P1e", float(np.max(np.abs(e.eval_p1e(params[:1])[0] - d["point0/p1e"]))), 5e-12, failures)
            if len(dets) == 1 and "point0/perbin_CC_night" in d:
                pb = e.eval_perbin(params[:1])
                for k in ("ES_day", "ES_night", "CC_day", "CC_night"):
                    check(f"point 0 per-bin {k}", rel(pb[k][0], d[f"point0/perbin_{k}"]), 1e-9, failures)
    else:
        print("no CUDA device (or --no-gpu): engine not checked")
    if a.save:
        np.savez_compressed(a.save, **{k.replace("/", "__"): v for k, v in d.items()})
        print(f"fixture saved to {a.save}")
    print("PARITY PINNED" if not failures else f"{len(failures)} check(s) FAILED: {failures[:6]}")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
