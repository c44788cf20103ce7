#!/usr/bin/env python
"""Build the fixed input tables of the measurement configuration.

Runs ONLY in the build container, where the reference checkout (and therefore its
``inputs/`` directory) is mounted at /root/reference; the GPU box has no such
directory, so the result is committed as ``tests/golden/tables_escc.npz`` and
everything else (tests, bench.py, smoke) reads that file.

Real inputs where the reference ships them (PREM, solar model, fluxes, exposure,
MARLEY cross-section, CC response/efficiency MC, CC background MC); synthetic
stand-ins, generated analytically, for the ES response / efficiency / background
MC that the reference lists in ``.MISSING_LARGE_BLOBS``.

    python tests/golden/make_tables.py [--ref /root/reference] [--out tests/golden/tables_escc.npz]
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
from scipy.special import erf
from scipy.stats import truncnorm

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from dune_solaroscfitter_b200.host import hdf5_min, table_builders as tb  # noqa: E402


def _gauss_cdf(x, mu, sig):
    return 0.5 * (1.0 + erf((x - mu) / (sig * math.sqrt(2.0))))


def synthetic_es_response(etrue_centres, reco: tb.BinSpec, frac):
    """Row-normalised Gaussian smearing, mean frac*E, sigma 0.12*sqrt(E * 1 MeV) (GeV units)."""
    edges = np.linspace(reco.min, reco.max, reco.bin_number + 1)
    mu = frac * etrue_centres
    sig = 0.12 * np.sqrt(etrue_centres * 1e-3)
    m = _gauss_cdf(edges[None, 1:], mu[:, None], sig[:, None]) - _gauss_cdf(edges[None, :-1], mu[:, None], sig[:, None])
    rs = m.sum(axis=1, keepdims=True)
    return np.divide(m, rs, out=np.zeros_like(m), where=rs > 0)


def smoothstep(x, lo, hi):
    t = np.clip((x - lo) / (hi - lo), 0.0, 1.0)
    return t * t * (3 - 2 * t)


def build(ref: str, n_cosz: int = 50, n_cosz_fine: int = 20, cosz_fine_edge: float = -0.15,
          n_etrue: int = 40):
    inp = os.path.join(ref, "inputs")
    # ---- shared binning (config.yaml:76-88) -------------------------------------------------
    etrue = tb.BinSpec(n_etrue, 1e-3, 21e-3)
    es_bins = tb.BinSpec(30, 2e-3, 20e-3)      # VD block, config.yaml:132-137
    cc_bins = tb.BinSpec(30, 1e-3, 25e-3)      # HD block, config.yaml:168-173
    thr_es, thr_cc = 5e-3, 5e-3
    es_norm, cc_norm = 20.0, 10.0              # HD exposures, config.yaml:160-161

    # ---- Earth ----------------------------------------------------------------------------
    model = tb.load_earth_model(os.path.join(inp, "EARTH_MODEL_PREM_DENSE.dat"))
    earth = tb.Linear1D(model["radius"], model["potential"], mode="throw")
    ex_x, ex_y = tb.load_exposure(os.path.join(inp, "solarExposure.csv"))
    grid = tb.build_cosz_grid(model, n_cosz, ex_x, n_fine=n_cosz_fine, fine_edge=cosz_fine_edge)
    prior = hdf5_min.load(os.path.join(inp, "earth_uncertainty_cov.jld2"))
    n_layers = int(prior["true"].shape[0])
    paths, lookup = tb.build_earth_tables(earth, grid.cosz_calc, n_layers)
    seg_start, seg_layer, seg_length = tb.pack_paths(paths)
    w_z = tb.exposure_weights(ex_x, ex_y, grid.edges)

    # ---- Sun / flux / cross-sections ------------------------------------------------------------
    ssm = tb.solar_model(hdf5_min.load(os.path.join(inp, "AGSS09_high_z.jld2")))
    fl = hdf5_min.load(os.path.join(inp, "fluxes.jld2"))
    f8, fh = tb.normalised_fluxes(fl["energies"], fl["flux8B"], fl["fluxHep"], etrue.max)
    ccx = tb.load_cc_xsec(os.path.join(inp, "CC_nue_40Ar_total_marley.csv"))
    unosc = tb.build_unoscillated_sample(etrue, f8, fh, ccx, thr_es, es_norm, cc_norm)
    _, e_fine = tb.calculate_bins(tb.BinSpec(2 * n_etrue, etrue.min, etrue.max))  # setup.jl:39-41
    _, e_centres = tb.calculate_bins(etrue)

    # ---- CC response / efficiency / backgrounds from the shipped MC -----------------------------
    es_ext, cc_ext = tb.extended_reco_bins(es_bins), tb.extended_reco_bins(cc_bins)
    cc_mc = tb.read_csv_columns(os.path.join(inp, "sergio_CC_response.csv"))
    cc_eff_mc = tb.read_csv_columns(os.path.join(inp, "eff_CC_sergio.csv"))
    assert np.array_equal(cc_mc["Ereco"], cc_eff_mc["Ereco"]), "CC response and efficiency MC are row-aligned"
    resp_cc = tb.create_response_matrix(cc_mc["Enu"], cc_mc["Ereco"], etrue, cc_ext)
    eff_cc = tb.selection_efficiency(cc_mc["Ereco"], cc_eff_mc["Mask"], cc_ext)
    ti, ri, ok = tb.response_bin_indices(cc_mc["Enu"], cc_mc["Ereco"], etrue, cc_ext)
    raw_cc = []
    for name in ("neutron_bg_CC_sergio.csv", "gamma_bg_CC_sergio.csv"):
        d = tb.read_csv_columns(os.path.join(inp, name))
        raw_cc.append(tb.cc_background_component(d["Ereco"], d["weights"], None, cc_ext, cc_norm))
    raw_cc = np.array(raw_cc)

    # ---- ES: synthetic stand-ins -------------------------------------------
    resp_es_nue = synthetic_es_response(e_centres, es_bins, 0.55)
    resp_es_nuo = synthetic_es_response(e_centres, es_bins, 0.50)
    es_edges = np.linspace(es_bins.min, es_bins.max, es_bins.bin_number + 1)
    es_c = 0.5 * (es_edges[1:] + es_edges[:-1])
    eff_es = 0.6 * smoothstep(es_c, 3e-3, 6e-3)
    shape = np.exp(-es_c / 2e-3)
    shape /= shape.sum()
    raw_es = np.array([tb.es_background_component(shape, eff_es, n, es_norm) for n in (1e6, 3e5, 1e5)])
    sides = np.array([0, -1, 1], dtype=np.int32)

    # ---- Asimov truth (config.yaml:27-31, readConfig.jl:337) ----------------------------------------
    mu, sd, lo, hi = 5.25e6, 0.28e6, 1.5e6, 1e7
    phi8 = float(truncnorm.mean((lo - mu) / sd, (hi - mu) / sd, loc=mu, scale=sd))
    truth = np.array([0.307, 0.0224, 6e-5, phi8, 8250.0, 1.0, 0.0, 0.0]
                     + [1.0] * n_layers + [2.2e-6] * 3 + [2.2e-6] * 2)

    return dict(
        n_cosz_sub=np.int32(grid.n_sub), cosz_edges=grid.edges, cosz_calc=grid.cosz_calc,
        e_calc=e_fine, etrue_centres=e_centres,
        seg_start=seg_start, seg_layer=seg_layer, seg_length=seg_length, earth_lookup=lookup,
        exposure_weights=w_z,
        avg_ne_8b=np.float64(ssm["avgNeBoron"]), avg_ne_hep=np.float64(ssm["avgNeHep"]),
        solar_ne=ssm["n_e"], solar_prod_8b=ssm["prodFractionBoron"], solar_prod_hep=ssm["prodFractionHep"],
        unosc_es_nue_8b=unosc["ES_nue_8B"], unosc_es_nuother_8b=unosc["ES_nuother_8B"], unosc_cc_8b=unosc["CC_8B"],
        unosc_es_nue_hep=unosc["ES_nue_hep"], unosc_es_nuother_hep=unosc["ES_nuother_hep"], unosc_cc_hep=unosc["CC_hep"],
        log_e_norm=unosc["log_E_norm"],
        resp_es_nue=resp_es_nue, resp_es_nuother=resp_es_nuo, resp_cc=resp_cc,
        eff_es_nue=eff_es, eff_es_nuother=eff_es.copy(), eff_cc=eff_cc,
        raw_bg_es=raw_es, bg_es_sides=sides, bg_es_par=np.ones(3, dtype=np.int32),
        raw_bg_cc=raw_cc, bg_cc_par=np.ones(2, dtype=np.int32),
        index_es=np.int32(tb.threshold_index(es_ext, thr_es)), index_cc=np.int32(tb.threshold_index(cc_ext, thr_cc)),
        es_reco_edges=es_ext.bins, cc_reco_edges=cc_ext.bins,
        cc_xsec_norm_true=np.float64(1.0), truth=truth,
        cc_mc_true_bin=ti.astype(np.int16), cc_mc_reco_bin=ri.astype(np.int16), cc_mc_in_range=ok,
    )


def build_dense(ref: str, n_cosz: int, n_etrue: int):
    """The measurement configuration on a dense zenith x energy grid (SURVEY.md section 8(d) C4;
    utils/checkOscGridConvergence.jl:36-46,122-145 sweeps 1000 cos z x 1000 E): the same inputs, detector tables
    included -- the CC response rebinned from the shipped MC to n_etrue true-energy bins, the unoscillated spectra
    integrated on them -- without the per-event bin indices."""
    t = build(ref, n_cosz=n_cosz, n_cosz_fine=0, n_etrue=n_etrue)
    for k in ("cc_mc_true_bin", "cc_mc_reco_bin", "cc_mc_in_range", "solar_ne", "solar_prod_8b", "solar_prod_hep"):
        t.pop(k)
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "tables_escc.npz"))
    a = ap.parse_args()
    t = build(a.ref)
    np.savez_compressed(a.out, **t)
    n_seg = int(t["seg_start"][-1])
    per = np.diff(t["seg_start"])
    print(f"paths={len(per)} segments={n_seg} max_per_path={per.max()} single={int((per == 1).sum())} "
          f"layers={len(t['earth_lookup'])} index_es={t['index_es']} index_cc={t['index_cc']}")
    print("earth_lookup", t["earth_lookup"])
    print("sum w_z", t["exposure_weights"].sum())
    here = os.path.dirname(a.out)
    for name, nz, nt in (("tables_dense200.npz", 200, 100), ("tables_dense1000.npz", 1000, 500)):
        d = build_dense(a.ref, nz, nt)
        np.savez_compressed(os.path.join(here, name), **d)
        per = np.diff(d["seg_start"])
        print(f"{name}: paths={len(per)} n_sub={d['n_cosz_sub']} segments={int(per.sum())} max={per.max()} energies={len(d['e_calc'])}")


if __name__ == "__main__":
    main()
