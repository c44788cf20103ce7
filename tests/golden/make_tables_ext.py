#!/usr/bin/env python
"""Extra tables of the inclusive / semi-inclusive / angular branches for the measurement configuration
(VD block of config.yaml:123-156: inclusiveMode, use_scattering_info, nBins_cos_scatter 25, ES_cos_cut 0).

Runs only where /root/reference is mounted (the CC MC is read from its inputs/); writes
tests/golden/tables_ext.npz, which tests merge over tables_escc.npz.  The ES angular MC
(inputs/ES_angular_reco.csv) is one of the reference's missing blobs, so ES.angular is an analytic
stand-in: per reco bin a forward-peaked exponential in cos(theta) on top of an isotropic floor.

    python tests/golden/make_tables_ext.py [--ref /root/reference]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from dune_solaroscfitter_b200.host import table_builders as tb  # noqa: E402


def build(ref: str, base: dict, n_cos: int = 25, cos_cut: float = 0.0):
    inp = os.path.join(ref, "inputs")
    n_etrue = len(base["e_calc"]) // 2
    etrue = tb.BinSpec(n_etrue, 1e-3, 21e-3)
    es_bins = tb.BinSpec(30, 2e-3, 20e-3)
    cos_bins = tb.BinSpec(n_cos, -1.0, 1.0)
    cc_mc = tb.read_csv_columns(os.path.join(inp, "sergio_CC_response.csv"))
    cc_eff_mc = tb.read_csv_columns(os.path.join(inp, "eff_CC_sergio.csv"))
    resp_cc_incl, eff_cc_incl = tb.cc_inclusive_tables(cc_mc["Enu"], cc_mc["Ereco"], cc_eff_mc["Mask"], etrue, es_bins)
    # analytic ES.angular: columns (reco bins) sum to one over the cos bins
    edges = np.linspace(-1.0, 1.0, n_cos + 1)
    c = 0.5 * (edges[:-1] + edges[1:])
    es_edges = np.linspace(es_bins.min, es_bins.max, es_bins.bin_number + 1)
    e_c = 0.5 * (es_edges[:-1] + es_edges[1:])
    width = 0.25 * np.sqrt(5e-3 / e_c)
    r_ang = np.exp((c[:, None] - 1.0) / width[None, :]) + 0.02
    r_ang /= r_ang.sum(axis=0, keepdims=True)
    r_bg, amask = tb.angular_bg_tables(r_ang.shape, cos_bins, cos_cut)
    mis_nue, mis_nuo, f_above, f_below = tb.misid_tables(base["resp_es_nue"], base["resp_es_nuother"], r_ang, amask,
                                                        base["resp_cc"], etrue, es_bins)
    return dict(n_cos=np.int32(n_cos), angular_cos_cut=np.float64(cos_cut), resp_cc_incl=resp_cc_incl, eff_cc_incl=eff_cc_incl,
                eff_es_nue_boosted=tb.hermite_boost(base["eff_es_nue"], es_bins),
                eff_es_nuother_boosted=tb.hermite_boost(base["eff_es_nuother"], es_bins),
                resp_es_angular=r_ang, resp_bg_angular=r_bg, angular_mask=amask,
                resp_cc_misid_nue=mis_nue, resp_cc_misid_nuother=mis_nuo,
                cc_f_above=np.float64(f_above), cc_f_below=np.float64(f_below))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    base = dict(np.load(os.path.join(here, "tables_escc.npz")))
    np.savez_compressed(os.path.join(here, "tables_ext.npz"), **build(a.ref, base))


if __name__ == "__main__":
    main()
