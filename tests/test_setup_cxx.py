"""Set-up stage in C++ (csrc/solarosc_setup.cuh behind solarosc_setup_*): bit-exact bin indices, histogramming
semantics, response matrices and Earth path tables, against the Python restatement (host/table_builders.py), the
committed golden tables and -- where /root/reference is mounted -- the reference's shipped MC files."""
import ctypes as C
import os

import numpy as np
import pytest

from dune_solaroscfitter_b200 import _lib
from dune_solaroscfitter_b200.host import table_builders as tb

REF = "/root/reference"
have_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "inputs")), reason="reference inputs not mounted")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def bin_index(lib, edges, v):
    edges, v = np.ascontiguousarray(edges, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(len(v), dtype=np.int32)
    assert lib.solarosc_setup_bin_index(_d(edges), len(edges), _d(v), len(v), out.ctypes.data_as(ip)) == 0
    return out


def histogram(lib, v, b: tb.BinSpec, weights=None, keep=None, normalise=True):
    v = np.ascontiguousarray(v, dtype=np.float64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    e = None if b.bins is None else np.ascontiguousarray(b.bins, dtype=np.float64)
    out = np.empty(b.bin_number)
    st = lib.solarosc_setup_histogram(_d(v), _d(w), None if k is None else k.ctypes.data_as(C.POINTER(C.c_uint8)), len(v),
                                      b.bin_number, b.min, b.max, _d(e), int(normalise), _d(out))
    assert st == 0
    return out


def response(lib, x, y, bx: tb.BinSpec, by: tb.BinSpec):
    x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    ex = None if bx.bins is None else np.ascontiguousarray(bx.bins, dtype=np.float64)
    ey = None if by.bins is None else np.ascontiguousarray(by.bins, dtype=np.float64)
    out = np.empty((bx.bin_number, by.bin_number))
    xi, yi = np.empty(len(x), dtype=np.int32), np.empty(len(x), dtype=np.int32)
    st = lib.solarosc_setup_response_matrix(_d(x), _d(y), len(x), bx.bin_number, bx.min, bx.max, _d(ex), by.bin_number, by.min,
                                            by.max, _d(ey), _d(out), xi.ctypes.data_as(ip), yi.ctypes.data_as(ip))
    assert st == 0
    return out, xi, yi


def test_bin_index_is_searchsortedlast(lib):
    edges = np.array([-np.inf, 1.0, 2.0, 3.0, np.inf])
    v = np.array([-5.0, 1.0, np.nextafter(1.0, 0), 1.5, 2.0, 2.999, 3.0, 1e9])
    assert bin_index(lib, edges, v).tolist() == [1, 2, 1, 2, 3, 3, 4, 4]
    rng = np.random.default_rng(0)
    e = np.sort(rng.uniform(0, 1, 31))
    q = np.concatenate([rng.uniform(-0.1, 1.1, 20000), e, np.nextafter(e, -1), np.nextafter(e, 2)])
    assert np.array_equal(bin_index(lib, e, q), tb.bin_index_left_closed(e, q))


def test_histogram_semantics(lib):
    b = tb.BinSpec(3, 0.0, 3.0, np.array([0.0, 1.0, 2.0, 3.0]))
    assert histogram(lib, [0.0, 0.5, 1.0, 2.0, 2.5, 3.0, -1.0], b, normalise=False).tolist() == [2.0, 1.0, 2.0]
    w = np.array([1e16, 1.0, -1e16, 1.0])
    assert histogram(lib, np.zeros(4) + 0.5, b, weights=w, normalise=False)[0] == ((1e16 + 1.0) - 1e16) + 1.0   # file order
    rng = np.random.default_rng(1)
    v, wt, keep = rng.uniform(-1, 4, 5000), rng.uniform(0, 2, 5000), rng.uniform(0, 1, 5000) < 0.7
    b2 = tb.BinSpec(7, 0.0, 3.5)
    assert np.array_equal(histogram(lib, v, b2, wt, keep, normalise=False), tb.create_histogram(v[keep], b2, wt[keep], normalise=False))
    assert np.allclose(histogram(lib, v, b2), tb.create_histogram(v, b2), rtol=1e-15)


def test_response_matrix_range_quirk_and_random(lib):
    bx = tb.BinSpec(2, 0.0, 2.0)
    by = tb.extended_reco_bins(tb.BinSpec(2, 0.0, 2.0))
    x = np.array([0.5, 0.5, 1.5, 1.5, 2.0])
    y = np.array([0.5, 5.0, 1.5, -1.0, 2.0])
    m, xi, yi = response(lib, x, y, bx, by)
    assert m.tolist() == [[1.0, 0.0], [0.0, 1.0]] and xi.tolist() == [1, 0, 2, 0, 2] and yi.tolist() == [1, 0, 2, 0, 2]
    rng = np.random.default_rng(2)
    x, y = rng.uniform(0, 22e-3, 30000), rng.uniform(0, 30e-3, 30000)
    x[:5] = 0.0                                                     # exact zeros are nudged to 1e-9 (response.jl:19-24)
    bx, by = tb.BinSpec(40, 1e-3, 21e-3), tb.extended_reco_bins(tb.BinSpec(30, 1e-3, 25e-3))
    m, xi, yi = response(lib, x, y, bx, by)
    ti, ri, ok = tb.response_bin_indices(x, y, bx, by)
    assert np.array_equal(xi, np.where(ok, ti, 0)) and np.array_equal(yi, np.where(ok, ri, 0))
    assert np.array_equal(m, tb.create_response_matrix(x, y, bx, by))


@have_ref
def test_cc_response_from_the_shipped_mc_matches_golden(lib, tables):
    inp = os.path.join(REF, "inputs")
    mc = tb.read_csv_columns(os.path.join(inp, "sergio_CC_response.csv"))
    eff = tb.read_csv_columns(os.path.join(inp, "eff_CC_sergio.csv"))
    etrue, cc_ext = tb.BinSpec(40, 1e-3, 21e-3), tb.extended_reco_bins(tb.BinSpec(30, 1e-3, 25e-3))
    m, xi, yi = response(lib, mc["Enu"], mc["Ereco"], etrue, cc_ext)
    ok = tables["cc_mc_in_range"]
    assert np.array_equal(xi[ok], tables["cc_mc_true_bin"][ok]) and np.array_equal(yi[ok], tables["cc_mc_reco_bin"][ok])   # bit-exact bins
    assert np.array_equal(xi == 0, ~ok)
    assert np.array_equal(m, tables["resp_cc"])
    sel = histogram(lib, mc["Ereco"], cc_ext, keep=eff["Mask"], normalise=False)
    tot = histogram(lib, mc["Ereco"], cc_ext, normalise=False)
    assert np.array_equal(np.where(tot == 0, 0.0, sel / np.where(tot == 0, 1.0, tot)), tables["eff_cc"])
    bg = tb.read_csv_columns(os.path.join(inp, "neutron_bg_CC_sergio.csv"))
    assert np.array_equal(histogram(lib, bg["Ereco"], cc_ext, weights=bg["weights"], normalise=False) * 10 * 10.0, tables["raw_bg_cc"][0])


@have_ref
def test_earth_paths_match_golden(lib, tables):
    model = tb.load_earth_model(os.path.join(REF, "inputs", "EARTH_MODEL_PREM_DENSE.dat"))
    r, p = np.ascontiguousarray(model["radius"]), np.ascontiguousarray(model["potential"])
    cz = np.ascontiguousarray(tables["cosz_calc"], dtype=np.float64)
    cap = 16 * len(cz)
    ss, sl, ln = np.empty(len(cz) + 1, dtype=np.int32), np.empty(cap, dtype=np.int32), np.empty(cap)
    lookup, nl = np.zeros(8), C.c_int32(0)
    st = lib.solarosc_setup_earth_paths(_d(r), _d(p), len(r), _d(cz), len(cz), 6, ss.ctypes.data_as(ip), sl.ctypes.data_as(ip),
                                        _d(ln), cap, _d(lookup), C.byref(nl))
    assert st == 0 and nl.value == 6
    n = ss[-1]
    assert np.array_equal(ss, tables["seg_start"]) and np.array_equal(sl[:n], tables["seg_layer"])      # integer tables exact
    assert np.max(np.abs(ln[:n] - tables["seg_length"]) / tables["seg_length"]) < 1e-14
    assert np.max(np.abs(lookup[:6] - tables["earth_lookup"]) / tables["earth_lookup"]) < 1e-14
    # a wrong layer count is an error, not a silently different table
    assert lib.solarosc_setup_earth_paths(_d(r), _d(p), len(r), _d(cz), len(cz), 5, ss.ctypes.data_as(ip), sl.ctypes.data_as(ip),
                                          _d(ln), cap, _d(lookup), C.byref(nl)) == 1


@pytest.mark.gpu
def test_device_bin_index_equals_host(lib, tables):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    rng = np.random.default_rng(3)
    e = np.concatenate([[-np.inf], np.linspace(1e-3, 25e-3, 31)[1:-1], [np.inf]])
    v = np.ascontiguousarray(np.concatenate([rng.uniform(-1e-3, 30e-3, 1_000_000), e[1:-1], np.nextafter(e[1:-1], 0)]))
    out = np.empty(len(v), dtype=np.int32)
    with SolarOscEngine(tables) as eng:
        assert lib.solarosc_setup_bin_index_device(eng._h, _d(e), len(e), _d(v), len(v), out.ctypes.data_as(ip)) == 0
    assert np.array_equal(out, bin_index(lib, e, v))
