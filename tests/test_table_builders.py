"""Set-up stage restated in Python (host/table_builders.py): histogramming semantics, bit-exact bin
indices, path tables.  Parts that need the reference's input files run only where /root/reference is
mounted (the build container); the rest uses the committed golden tables."""
import os

import numpy as np
import pytest

from dune_solaroscfitter_b200.host import hdf5_min, table_builders as tb

REF = "/root/reference"
have_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "inputs")), reason="reference inputs not mounted")


def test_left_closed_bin_index_matches_searchsortedlast():
    edges = np.array([-np.inf, 1.0, 2.0, 3.0, np.inf])
    v = np.array([-5.0, 1.0, np.nextafter(1.0, 0), 1.5, 2.0, 2.999, 3.0, 1e9])
    assert tb.bin_index_left_closed(edges, v).tolist() == [1, 2, 1, 2, 3, 3, 4, 4]


def test_create_histogram_is_left_closed_and_ordered():
    b = tb.BinSpec(3, 0.0, 3.0, np.array([0.0, 1.0, 2.0, 3.0]))
    h = tb.create_histogram(np.array([0.0, 0.5, 1.0, 2.0, 2.5, 3.0, -1.0]), b, normalise=False)
    assert h.tolist() == [2.0, 1.0, 2.0]                 # 3.0 and -1.0 fall outside (histHelpers.jl:60-63)
    w = np.array([1e16, 1.0, -1e16, 1.0])
    hw = tb.create_histogram(np.zeros(4) + 0.5, b, weights=w, normalise=False)
    assert hw[0] == ((1e16 + 1.0) - 1e16) + 1.0          # weights summed in file order


def test_response_matrix_range_quirk():
    # in-range test on the UN-extended range although the bins run to +-Inf (response.jl:47 vs :98-104)
    bx = tb.BinSpec(2, 0.0, 2.0)
    by = tb.extended_reco_bins(tb.BinSpec(2, 0.0, 2.0))
    x = np.array([0.5, 0.5, 1.5, 1.5, 2.0])
    y = np.array([0.5, 5.0, 1.5, -1.0, 2.0])             # 5.0 and -1.0 are dropped, 2.0 lands in the last bin
    m = tb.create_response_matrix(x, y, bx, by)
    assert m.tolist() == [[1.0, 0.0], [0.0, 1.0]]
    ti, ri, ok = tb.response_bin_indices(x, y, bx, by)
    assert ok.tolist() == [True, False, True, False, True] and ti[-1] == 2 and ri[-1] == 2


def test_threshold_index_convention(tables):
    # findfirst(edge > thr) on the extended edges, used directly as first bin (setup.jl:169-178)
    ext = tb.BinSpec(30, 2e-3, 20e-3, tables["es_reco_edges"])
    k = tb.threshold_index(ext, 5e-3)
    assert k == int(tables["index_es"]) and tables["es_reco_edges"][k - 1] > 5e-3 >= tables["es_reco_edges"][k - 2]


def test_cosz_grid_allocation():
    c = tb._alloc_cosz_bins(30, np.array([0.0973, 0.6868]))
    assert c.tolist() == [4, 26] and c.sum() == 30
    with pytest.raises(ValueError):
        tb._alloc_cosz_bins(1, np.array([0.5, 0.5]))


def test_golden_tables_structure(tables):
    per = np.diff(tables["seg_start"])
    assert len(per) == 50 and per.sum() == 194 and per.max() == 9 and int((per == 1).sum()) == 24
    assert len(tables["earth_lookup"]) == 6 and np.all(np.diff(tables["earth_lookup"]) > 0)
    for p in range(50):      # layer sequences are palindromes
        seq = tables["seg_layer"][tables["seg_start"][p]:tables["seg_start"][p + 1]]
        assert np.array_equal(seq, seq[::-1])
    assert abs(tables["exposure_weights"].sum() - 1.0) < 1e-12
    rs = tables["resp_cc"].sum(axis=1)
    assert np.all((np.abs(rs - 1) < 1e-12) | (rs == 0))
    assert np.all(np.diff(tables["cosz_edges"]) > 0) and len(tables["cosz_edges"]) == 51


@have_ref
def test_bin_indices_are_bit_exact_against_the_shipped_mc(tables):
    mc = tb.read_csv_columns(os.path.join(REF, "inputs", "sergio_CC_response.csv"))
    etrue = tb.BinSpec(40, 1e-3, 21e-3)
    ext = tb.extended_reco_bins(tb.BinSpec(30, 1e-3, 25e-3))
    ti, ri, ok = tb.response_bin_indices(mc["Enu"], mc["Ereco"], etrue, ext)
    assert len(ti) == 54710
    assert np.array_equal(ti.astype(np.int16), tables["cc_mc_true_bin"])
    assert np.array_equal(ri.astype(np.int16), tables["cc_mc_reco_bin"])
    assert np.array_equal(ok, tables["cc_mc_in_range"])
    # no event sits within 4 ulp of an edge, so last-ulp differences between numpy.linspace and Julia's
    # TwicePrecision range cannot move an event across a bin boundary (hard part 6)
    for vals, edges in ((mc["Enu"], np.linspace(1e-3, 21e-3, 41)), (mc["Ereco"], np.linspace(1e-3, 25e-3, 31))):
        d = np.min(np.abs(vals[:, None] - edges[None, :]) / np.spacing(edges)[None, :], axis=1)
        assert d.min() > 4


@have_ref
def test_regenerated_tables_equal_the_golden_file(tables):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_tables", os.path.join(os.path.dirname(__file__), "golden", "make_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.build(REF)
    for k, v in fresh.items():
        assert np.array_equal(np.asarray(v), tables[k]), k


@have_ref
def test_hdf5_reader_on_reference_inputs():
    d = hdf5_min.load(os.path.join(REF, "inputs", "fluxes.jld2"))
    assert d["energies"].shape == (999,) and float(d["total8B"]) == 4.5e6 and float(d["totalHep"]) == 8250.0
    c = hdf5_min.load(os.path.join(REF, "inputs", "earth_uncertainty_cov.jld2"))
    assert np.array_equal(c["covariance"], 0.0144 * np.eye(6)) and np.array_equal(c["true"], np.ones(6))
