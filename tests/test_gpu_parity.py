"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on identical tables.

Tolerances are BASELINE.json's: per-bin rates <= 1e-9 relative, -2lnL <= 1e-6 absolute.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RATE_RTOL = 1e-9
M2LNL_ATOL = 1e-6


@pytest.fixture(scope="module")
def engine(tables, asimov):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    with SolarOscEngine(tables, data=asimov) as e:
        yield e


@pytest.fixture(scope="module")
def coracle(tables, asimov):
    from oracle import oracle_c
    return oracle_c.COracle(tables, asimov)


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_p1e_matches_oracle(engine, coracle, tables, draw):
    p = draw(tables, 6, wide=True)
    p[0] = tables["truth"]
    got = engine.eval_p1e(p)
    _, ref = coracle.eval_oscprobs(p, want_p1e=True)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 5e-12


def test_vacuum_limit(engine, tables):
    p = np.array(tables["truth"], dtype=np.float64)
    p[8:14] = 0.0     # zero density in every layer: P1e = c12^2 c13^2 and night == day
    p1e = engine.eval_p1e(p)
    assert np.max(np.abs(p1e - (1 - p[0]) * (1 - p[1]))) < 1e-14
    o = engine.eval_oscprobs(p)
    assert np.max(np.abs(o["nue_8B_night"][0] - o["nue_8B_day"][0][None, :])) < 1e-14


def test_oscprobs_match_oracle(engine, coracle, tables, draw):
    p = draw(tables, 8)
    got = engine.eval_oscprobs(p)
    ref = coracle.eval_oscprobs(p)
    assert np.max(np.abs(got["nue_8B_day"] - ref[:, 0, 0])) < 1e-14
    assert np.max(np.abs(got["nue_hep_day"] - ref[:, 1, 0])) < 1e-14
    assert np.max(np.abs(got["nue_8B_night"] - ref[:, 0, 1:])) < 5e-12
    assert np.max(np.abs(got["nue_hep_night"] - ref[:, 1, 1:])) < 5e-12


def test_rates_match_oracle(engine, coracle, tables, draw):
    p = draw(tables, 16)
    p[0] = tables["truth"]
    got = engine.eval_rates(p)
    ref = coracle.eval_rates(p)
    for k in ("ES_day", "ES_night", "CC_day", "CC_night", "BG_ES_tot", "BG_CC_tot"):
        assert got[k].shape == ref[k].shape, k
        assert _rel(got[k], ref[k]) < RATE_RTOL, (k, _rel(got[k], ref[k]))


def test_llh_matches_oracle(engine, coracle, tables, draw):
    p = draw(tables, 96)
    p[0] = tables["truth"]
    got, fl = engine.eval_llh(p, return_flags=True)
    ref, rfl = coracle.eval_llh(p)
    assert np.array_equal(fl, rfl)
    assert np.max(np.abs(-2 * got - (-2 * ref))) < M2LNL_ATOL
    assert abs(got[0]) < 1e-6    # Asimov: lnL(truth) == 0 up to rounding


def test_llh_numpy_twin_spot_check(engine, tables, asimov, draw):
    from oracle import oracle_np
    p = draw(tables, 3, seed=7)
    got = engine.eval_llh(p)
    ref = np.array([oracle_np.total_llh(q, tables, asimov) for q in p])
    assert np.max(np.abs(2 * got - 2 * ref)) < M2LNL_ATOL


def test_earth_norm_bounds_give_minus_inf(engine, tables):
    p = np.tile(np.asarray(tables["truth"], dtype=np.float64), (3, 1))
    p[1, 8] = -0.1
    p[2, 10] = 2.1
    got, fl = engine.eval_llh(p, return_flags=True)
    assert np.isfinite(got[0]) and got[1] == -np.inf and got[2] == -np.inf
    assert fl[1] == 1 and fl[2] == 1


def test_nan_parameters_give_flagged_nan(engine, tables):
    """A NaN in any column reaches the result as NaN with the NONFINITE flag (Julia's clamp, asin, exp and the Poisson
    sum all propagate it); a NaN earth norm fails the bounds check like any other out-of-range value: -Inf."""
    truth = np.asarray(tables["truth"], dtype=float)
    p = np.tile(truth, (20, 1))
    for c in range(19):
        p[c, c] = np.nan
    for batch in (p, np.tile(p, (10, 1))):            # small batch: osc_kernel + rates_kernel; large: fused llh_kernel
        got, fl = engine.eval_llh(batch, return_flags=True)
        got, fl = got[:20], fl[:20]
        for c in range(19):
            if 8 <= c < 14:
                assert got[c] == -np.inf and fl[c] == 1, c
            elif c in (15, 16):                        # ES_bg_norm_2, _3 are never read (propagation_bg.jl:56-57)
                assert abs(got[c]) < 1e-6 and fl[c] == 0, c
            else:
                assert np.isnan(got[c]) and fl[c] & 8, c
        assert abs(got[19]) < 1e-6 and fl[19] == 0


def test_grouped_equals_ungrouped(engine, tables, draw):
    g = 9
    p = draw(tables, 4 * g)
    for c in range(4):    # a scan cell: oscillation columns constant, nuisances vary
        p[c * g:(c + 1) * g, [0, 1, 2]] = p[c * g, [0, 1, 2]]
        p[c * g:(c + 1) * g, 8:14] = p[c * g, 8:14]
    a = engine.eval_llh(p)
    b = engine.eval_llh(p, group=g)
    # same probabilities, same Poisson terms; the grouped fold cuts its work into finer items, so the terms are
    # summed in a different (still fixed) order
    assert np.max(np.abs(a - b)) < 1e-9
    assert np.array_equal(b, engine.eval_llh(p, group=g))      # each arrangement is deterministic


def test_chunking_and_determinism(tables, asimov, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    p = draw(tables, 300)
    with SolarOscEngine(tables, data=asimov, max_batch=64) as small, SolarOscEngine(tables, data=asimov) as big:
        a, b, c = small.eval_llh(p), big.eval_llh(p), big.eval_llh(p)
    assert np.array_equal(b, c)
    # different batch sizes change the energy-chunking of the oscillation kernel but not the arithmetic
    assert np.array_equal(a, b)


def test_small_batches_use_energy_chunks(engine, coracle, tables, draw):
    for n in (1, 2, 10):
        p = draw(tables, n, seed=n)
        ref, _ = coracle.eval_llh(p)
        assert np.max(np.abs(2 * engine.eval_llh(p) - 2 * ref)) < M2LNL_ATOL


def test_perbin_matches_oracle(engine, tables, asimov, draw):
    from oracle import oracle_np
    p = draw(tables, 2, seed=3)
    p[1, 9] = 2.5
    got = engine.eval_perbin(p)
    for i in range(2):
        ref = oracle_np.perbin_llh(p[i], tables, asimov)
        for k in ("ES_day", "ES_night", "CC_day", "CC_night"):
            if i == 1:
                assert np.all(got[k][i] == -np.inf)
            else:
                scale = np.maximum(np.abs(ref[k]), 1.0)
                assert np.max(np.abs(got[k][i] - ref[k]) / scale) < 1e-9, k


def test_slow_mode_matches_oracle(tables, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_c, oracle_np
    data = oracle_np.asimov_data(tables, fast=False)
    p = draw(tables, 4)
    ref_o = oracle_c.COracle(tables, data, fast=False)
    with SolarOscEngine(tables, data=data, fast=False) as e:
        got = e.eval_oscprobs(p)
        ref = ref_o.eval_oscprobs(p)
        assert np.max(np.abs(got["nue_8B_day"] - ref[:, 0, 0])) < 1e-13
        assert np.max(np.abs(got["nue_8B_night"] - ref[:, 0, 1:])) < 5e-12
        lr, _ = ref_o.eval_llh(p)
        assert np.max(np.abs(2 * e.eval_llh(p) - 2 * lr)) < M2LNL_ATOL


def test_col_major_layout(tables, asimov, draw):
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    p = draw(tables, 3)
    t2 = dict(tables)
    for k in ("resp_es_nue", "resp_es_nuother", "resp_cc"):
        t2[k] = np.ascontiguousarray(tables[k].T)   # a numpy view of Julia's column-major memory
    d2 = {k: (np.ascontiguousarray(v.T) if v.ndim == 2 else v) for k, v in asimov.items()}
    with SolarOscEngine(tables, data=asimov) as a, SolarOscEngine(t2, data=d2, col_major=True) as b:
        ra, rb = a.eval_rates(p), b.eval_rates(p)
        assert np.array_equal(a.eval_llh(p), b.eval_llh(p))
        assert np.array_equal(ra["ES_night"], rb["ES_night"].transpose(0, 2, 1))
        assert np.array_equal(ra["CC_night"], rb["CC_night"].transpose(0, 2, 1))
        assert np.array_equal(a.eval_p1e(p), b.eval_p1e(p).transpose(0, 2, 1))


def test_poisson_kats(engine):
    # reference tests/test_likelihood_correctness.jl:5-17
    e = np.array([10.0, 20.0, 30.0, 15.0]); m = np.array([12.0, 18.0, 29.0, 14.0])
    ref = np.sum((e - m) + m * np.log(m / e))
    assert abs(engine.poisson_llh(e, m)[0] + ref) < 1e-10
    n = np.array([1.0, 5.0, 10.0])
    assert abs(engine.poisson_llh(n, n)[0]) < 1e-12
    assert np.isfinite(engine.poisson_llh([0.0, 1.0, 2.0], [0.0, 0.0, 3.0])[0])
    with pytest.raises(ValueError):
        engine.poisson_llh([1.0, 2.0], [1.0])


def test_poisson_benchmark_shape(engine):
    # reference tests/benchmark_likelihoods.jl:10-19: two length-10000 vectors of U{10..50}
    from oracle import oracle_c
    rng = np.random.default_rng(5)
    e = rng.integers(10, 51, (4, 10000)).astype(float); m = rng.integers(10, 51, (4, 10000)).astype(float)
    got = engine.poisson_llh(e, m)
    ref = np.array([oracle_c.COracle.poisson(e[i], m[i])[0] for i in range(4)])
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12


def test_negative_rate_flags(tables, asimov):
    from dune_solaroscfitter_b200 import abi
    from dune_solaroscfitter_b200.engine import NegativeRateError, SolarOscEngine
    p = np.array(tables["truth"], dtype=np.float64)
    p[-5:] = -2.2e-6    # negative background norms: expected rates far below -1e-3
    with SolarOscEngine(tables, data=asimov) as e:
        got, fl = e.eval_llh(p, return_flags=True)
        assert np.isnan(got[0]) and fl[0] & abi.FLAG_NEG_ERROR
        with pytest.raises(NegativeRateError):
            e.eval_llh(p, raise_on_negative=True)


def test_device_resident_entry(engine, tables, draw):
    torch = pytest.importorskip("torch")
    p = draw(tables, 513)
    ref = engine.eval_llh(p)
    dp = torch.from_numpy(p).cuda()
    out = torch.empty(len(p), dtype=torch.float64, device="cuda")
    fl = torch.zeros(len(p), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    engine.eval_llh_device(dp.data_ptr(), len(p), out.data_ptr(), fl.data_ptr(),
                           stream_ptr=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)


def test_parameter_corners(engine, coracle, tables):
    """Degenerate and extreme parameter points (the cases of tests/test_host_math.py plus flux / cross-section /
    background extremes) through the full GPU path: P1e, rates and -2lnL against the oracle."""
    truth = np.asarray(tables["truth"], dtype=float)
    corners = [
        dict(c0=0.454, c1=0.0, c2=8.6e-5),                      # theta13 = 0: the third state decouples exactly
        dict(c0=0.0, c1=0.0232, c2=1.4e-4),                     # theta12 = 0
        dict(c0=1.0, c1=0.0095, c2=1.5e-4),                     # theta12 = pi/2
        dict(c0=0.32, c1=0.0317, c2=2e-5, norms=2.0),           # strongest matter effect inside the priors
        dict(c0=0.462, c1=1e-12, c2=1.9e-4),                    # shift far below the rounding noise of the cubic
        dict(c0=0.4, c1=0.25, c2=4e-5),                         # llhScan.jl's theta13 upper limit
        dict(c0=0.2, c1=1e-4, c2=2.08e-4, norms=0.5),
        dict(c0=0.307, c1=0.0224, c2=6e-5, norms=0.0),          # vacuum: all earth norms 0
        dict(c0=0.307, c1=0.0224, c2=1e-6),                     # dm21 far below the matter potential
        dict(c0=0.307, c1=0.0224, c2=1e-3),                     # dm21 close to dm31
        dict(c0=1.3, c1=-0.2, c2=6e-5),                         # out-of-range mixings are clamped (propagation_osc.jl:30-36)
        dict(c0=0.307, c1=0.0224, c2=6e-5, c3=0.0, c4=0.0),     # no signal at all: backgrounds only
        dict(c0=0.307, c1=0.0224, c2=6e-5, c5=2.0, c6=0.5, c7=-0.3),
        dict(c0=0.307, c1=0.0224, c2=6e-5, c14=0.0, c17=0.0, c18=1e-5),
    ]
    p = np.tile(truth, (len(corners), 1))
    for i, c in enumerate(corners):
        for k, v in c.items():
            if k == "norms":
                p[i, 8:14] = v
            else:
                p[i, int(k[1:])] = v
    got, fl = engine.eval_llh(p, return_flags=True)
    ref, rfl = coracle.eval_llh(p)
    assert np.array_equal(fl, rfl)
    assert np.max(np.abs(2 * got - 2 * ref)) < 1e-6
    big, bfl = engine.eval_llh(np.tile(p, (12, 1)), return_flags=True)   # the fused kernel, against the oracle as well
    assert engine.last_timing()["fused"]
    assert np.max(np.abs(2 * big.reshape(12, -1) - 2 * ref[None, :])) < 1e-6
    assert np.array_equal(bfl.reshape(12, -1), np.tile(rfl, (12, 1)))
    assert np.max(np.abs(big.reshape(12, -1) - got[None, :])) < 1e-7
    p1e = engine.eval_p1e(p)
    _, rp1e = coracle.eval_oscprobs(p, want_p1e=True)
    assert np.max(np.abs(p1e - rp1e)) < 5e-12
    r, rr = engine.eval_rates(p), coracle.eval_rates(p)
    for k in ("ES_day", "ES_night", "CC_day", "CC_night"):
        scale = np.maximum(np.abs(rr[k]), 1e-9 * np.max(np.abs(rr[k])))
        assert np.max(np.abs(r[k] - rr[k]) / scale) < 1e-9, k


def test_mirrored_walk_equals_the_generic_walk_on_the_device(tables, asimov, draw, monkeypatch):
    """SOLAROSC_NO_MIRROR makes solarosc_create flag no path as mirrored: every chord then takes the generic step-by-step
    loop.  P1e of the two walks agrees to rounding (a few 1e-16), lnL to the Poisson sum's noise."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    p = draw(tables, 600, seed=41, wide=True)
    with SolarOscEngine(tables, data=asimov) as e:
        a_p1e, a_lnl, a_small = e.eval_p1e(p[:8]), e.eval_llh(p), e.eval_llh(p[:8])
    monkeypatch.setenv("SOLAROSC_NO_MIRROR", "1")
    with SolarOscEngine(tables, data=asimov) as e:
        b_p1e, b_lnl, b_small = e.eval_p1e(p[:8]), e.eval_llh(p), e.eval_llh(p[:8])
    assert np.max(np.abs(a_p1e - b_p1e)) < 5e-15 and not np.array_equal(a_p1e, b_p1e)    # (two different instruction streams)
    assert np.max(np.abs(2 * a_lnl - 2 * b_lnl)) < 1e-7 and np.max(np.abs(2 * a_small - 2 * b_small)) < 1e-7


@pytest.mark.parametrize("factor", [10.0, 100.0])
def test_exposure_scaled_parity(tables, draw, factor):
    """The rounding error of -2lnL grows with the counts per bin: the reference forms m (log m - log e) from two rounded
    logarithms (likelihood_statistical.jl:19-23), so ANY two faithful double-precision evaluations -- Julia's log, glibc's,
    the kernel's log_pos -- differ by about m ulp(log m) per bin, 3e-8 for the 1.8e7-count CC background bins of the
    nominal exposure.  Exposure x 10 and x 100 (signal and background tables scaled, Asimov data regenerated), the fused
    kernel against the oracle on prior draws: the error stays proportional to the exposure (measured 2.2e-7, 2.2e-6,
    2.4e-5 at x 1, 10, 100), i.e. the 1e-6 tolerance of the nominal exposure scales with the counts, and stays far below
    the worst case of the formula's own rounding, sum_bins 2 m ulp(log m)."""
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_c, oracle_np
    t = dict(tables)
    for k in ("unosc_es_nue_8b", "unosc_es_nuother_8b", "unosc_es_nue_hep", "unosc_es_nuother_hep", "unosc_cc_8b", "unosc_cc_hep",
              "raw_bg_es", "raw_bg_cc"):
        t[k] = factor * np.asarray(tables[k])
    data = oracle_np.asimov_data(t)
    p = draw(t, 1024, seed=77)
    p[0] = t["truth"]
    ref, _ = oracle_c.COracle(t, data).eval_llh(p)
    with SolarOscEngine(t, data=data) as e:
        got = e.eval_llh(p)
        assert e.last_timing()["fused"]
    err = np.abs(2 * got - 2 * ref)
    ie, ic = int(t["index_es"]) - 1, int(t["index_cc"]) - 1
    m = np.concatenate([np.ravel(data["ES_day"][ie:]), np.ravel(data["ES_night"][:, ie:]), np.ravel(data["CC_day"][ic:]),
                        np.ravel(data["CC_night"][:, ic:])])
    m = m[m > 0]
    formula_noise = 2 * float(np.sum(2 * m * np.spacing(np.abs(np.log(m)))))     # -2lnL: two logarithms, one ulp each, every bin
    print(f"exposure x{factor:g}: max |d(-2lnL)| = {err.max():.2e} (tolerance {1e-6 * factor:.0e}); worst case of the formula's own "
          f"rounding {formula_noise:.1e}; -2lnL up to {np.abs(2 * ref).max():.2e}")
    assert abs(got[0]) < 1e-6 * factor
    assert err.max() <= 1e-6 * factor
    assert err.max() < 0.5 * formula_noise


def test_single_rank_communicator_gathers_the_same_values(tables, asimov, draw):
    """The in-library NCCL path with a one-rank communicator: solarosc_eval_llh_sharded == solarosc_eval_llh (the
    multi-rank run is tools/sharded_check.py under torchrun)."""
    import ctypes as C
    from dune_solaroscfitter_b200 import abi
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    p = np.ascontiguousarray(draw(tables, 27 * 40, seed=91))
    p[5, 9] = 2.5
    with SolarOscEngine(tables, data=asimov) as e:
        e.comm_init(e.comm_unique_id(), 0, 1)
        assert e.shard_bounds(len(p), 27) == (0, len(p))
        a, fa = e.eval_llh_sharded(p, len(p), return_flags=True)
        b, fb = e.eval_llh(p, return_flags=True)
        g = e.eval_llh_sharded(p, len(p), group=27)
        gb = e.eval_llh(p, group=27)
        # a negative-rate error anywhere in the gathered batch comes back as the call's status (one word OR-ed on the
        # device), caller-owned result arrays
        q = p.copy(); q[700, -5:] = -2.2e-6
        out, fl = np.empty(len(q)), np.empty(len(q), dtype=np.uint32)
        st = e._L.solarosc_eval_llh_sharded(e._h, q.ctypes.data_as(C.POINTER(C.c_double)), len(q), 1,
                                            out.ctypes.data_as(C.POINTER(C.c_double)), fl.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert st == abi.ERR_NEGATIVE_RATE and np.isnan(out[700]) and fl[700] & abi.FLAG_NEG_ERROR
        assert e._L.solarosc_eval_llh_sharded(e._h, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), 1,
                                              out.ctypes.data_as(C.POINTER(C.c_double)), None) == abi.OK
        assert np.array_equal(out, b)
    assert np.array_equal(a, b) and np.array_equal(fa, fb) and np.array_equal(g, gb)
