"""The kernels' arithmetic helpers (csrc/solarosc_device.cuh), compiled for the host by nvcc and run on the
CPU against the oracle: sin/cos reduction, rank-one eigen-system, layer-step algebra.  This is a test
harness (tests/csrc/host_math_check.cu); the library never runs these on the CPU."""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "host_math_check.cu")
OUT = os.path.join(HERE, "csrc", "_build", "libhostmath.so")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def hm():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    dev = os.path.join(HERE, "..", "dune_solaroscfitter_b200", "csrc", "solarosc_device.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(dev)):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-mfma",
                        "-shared", "-o", OUT, SRC], check=True, capture_output=True)
    L = C.CDLL(OUT)
    L.hm_sincos.argtypes = [dp, C.c_int, dp, dp]
    L.hm_sincos_qt.argtypes = [dp, dp, C.c_int, dp, dp]
    L.hm_eig.argtypes = [C.c_double] * 6 + [dp, dp]
    L.hm_p1e.argtypes = [dp, C.c_int, ip, ip, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    L.hm_p1e_mirror.argtypes = [dp, C.c_int, ip, ip, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp]
    L.hm_p1e_mirror.restype = C.c_int
    L.hm_msw_c2m.argtypes = [C.c_double] * 5
    L.hm_msw_c2m.restype = C.c_double
    return L


def test_sincos_phase_accuracy(hm):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1e5, 1e5, 200000), rng.uniform(-4, 4, 20000),
                        [0.0, math.pi / 4, -math.pi / 4, math.pi / 2, 7e4, 1e6, -1e6]])
    s, c = np.empty_like(x), np.empty_like(x)
    hm.hm_sincos(x.ctypes.data_as(dp), len(x), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert np.max(np.abs(s - np.sin(x))) < 4e-16
    assert np.max(np.abs(c - np.cos(x))) < 4e-16


def test_sincos_quarter_turns_accuracy(hm):
    """sincos_qt reduces the exact product len * kq, so the reference reduces it exactly too (rationals)."""
    from fractions import Fraction
    rng = np.random.default_rng(1)
    n = 20000
    ln = np.concatenate([rng.uniform(1.0, 1.3e4, n), [0.0, 1.0, 2.0, 12742.0, 1.0, 3.0]])
    kq = np.concatenate([rng.uniform(0.0, 4.0, n // 2), rng.uniform(0.0, 1e-3, n - n // 2), [1.0, 0.5, 0.25, 3.5, 1e-300, 0.5]])
    ref_s, ref_c = np.empty_like(ln), np.empty_like(ln)
    for i in range(len(ln)):
        x = Fraction(float(ln[i])) * Fraction(float(kq[i]))
        k = round(x)
        f = float(x - k) * (math.pi / 2)
        sr, cr = math.sin(f), math.cos(f)
        ref_s[i], ref_c[i] = [(sr, cr), (cr, -sr), (-sr, -cr), (-cr, sr)][k % 4]
    s, c = np.empty_like(ln), np.empty_like(ln)
    hm.hm_sincos_qt(ln.ctypes.data_as(dp), kq.ctypes.data_as(dp), len(ln), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    # degree-11 minimax sine kernel: 3.4e-15; it serves as the cosine in odd quadrants
    assert np.max(np.abs(s - ref_s)) < 5e-15 and np.max(np.abs(c - ref_c)) < 5e-15


def _eig(hm, d21, d31, a, u):
    q, w = np.empty(9), np.empty(2)
    hm.hm_eig(d21, d31, a, u[0], u[1], u[2], q.ctypes.data_as(dp), w.ctypes.data_as(dp))
    return q.reshape(3, 3), w


@pytest.mark.parametrize("a", [0.0, 1e-30, 3e-7, 1e-5, 4e-5])
@pytest.mark.parametrize("s13sq", [0.0, 1e-12, 0.0224, 0.25])
def test_rank_one_eigensystem(hm, a, s13sq):
    d21, d31, s12sq = 6e-5, 2.5e-3, 0.307
    c13 = math.sqrt(1 - s13sq)
    u = np.array([math.sqrt(1 - s12sq) * c13, math.sqrt(s12sq) * c13, math.sqrt(s13sq)])
    q, w = _eig(hm, d21, d31, a, u)
    s = np.diag([0.0, d21, d31]) + a * np.outer(u, u)
    vals, vecs = np.linalg.eigh(s)
    assert abs(w[0] - (vals[1] - vals[0])) < 1e-17 and abs(w[1] - (vals[2] - vals[0])) < 1e-17
    assert np.max(np.abs(q.T @ q - np.eye(3))) < 1e-13                      # orthonormal columns
    assert np.max(np.abs(np.abs(q.T @ vecs) - np.eye(3))) < 1e-10          # same eigenvectors up to sign
    assert np.max(np.abs(q @ np.diag(vals) @ q.T - s)) < 1e-17


def _applied(t):
    ss = np.ascontiguousarray(t["seg_start"], dtype=np.int32)
    lay, ln = np.empty(ss[-1], dtype=np.int32), np.empty(ss[-1])
    for p in range(len(ss) - 1):
        s0, n = ss[p], ss[p + 1] - ss[p]
        for m in range(n):    # applied step m: layer index[n-1-m], length[m] (osc.jl:436-446, 455-456)
            lay[s0 + m] = t["seg_layer"][s0 + n - 1 - m] - 1
            ln[s0 + m] = t["seg_length"][s0 + m]
    return ss, lay, ln


CASES = [
    (0.307, 0.0224, 6e-5, np.ones(6)),
    (0.261, 0.0255, 2.33e-5, np.array([0, 2, 0, 2, 1, 1.0])),
    (0.454, 0.0, 8.6e-5, np.ones(6)),              # theta13 = 0: third state decouples exactly
    (0.0, 0.0232, 1.4e-4, np.ones(6)),             # theta12 = 0
    (1.0, 0.0095, 1.5e-4, np.ones(6)),
    (0.32, 0.0317, 2e-5, np.full(6, 2.0)),         # strongest matter effect inside the priors
    (0.462, 1e-12, 1.9e-4, np.ones(6)),            # shift far below the rounding noise of the cubic
    (0.4, 0.25, 4e-5, np.ones(6)),                 # llhScan.jl's theta13 upper limit
    (0.2, 1e-4, 2.08e-4, np.full(6, 0.5)),
]


@pytest.mark.parametrize("s12,s13,dm,norms", CASES)
def test_kernel_algebra_matches_oracle_p1e(hm, tables, s12, s13, dm, norms):
    ss, lay, ln = _applied(tables)
    e = np.ascontiguousarray(tables["e_calc"])
    lookup = np.ascontiguousarray(norms * tables["earth_lookup"])
    got = np.empty((len(ss) - 1, len(e)))
    hm.hm_p1e(e.ctypes.data_as(dp), len(e), ss.ctypes.data_as(ip), lay.ctypes.data_as(ip), ln.ctypes.data_as(dp),
              len(ss) - 1, lookup.ctypes.data_as(dp), len(lookup), s12, s13, dm, got.ctypes.data_as(dp))
    mp = O.OscPars(dm, math.asin(math.sqrt(s12)), math.asin(math.sqrt(s13)))
    ref = O.osc_prob_earth(e, mp, lookup, tables["seg_start"], tables["seg_layer"], tables["seg_length"])
    assert np.max(np.abs(got - ref)) < 5e-12


@pytest.mark.parametrize("s12,s13,dm,norms", CASES)
def test_mirrored_pair_walk_matches_oracle_and_the_generic_walk(hm, tables, s12, s13, dm, norms):
    """path_p1e_mirror (psi from the nu_e end, chi from the nu_1 end, shared eigen entries, chi's phase factors by a
    second-order rotation of psi's) on every multi-layer path of the default binning: the oracle's P1e and, to rounding,
    the generic step-by-step walk."""
    ss, lay, ln = _applied(tables)
    e = np.ascontiguousarray(tables["e_calc"])
    lookup = np.ascontiguousarray(norms * tables["earth_lookup"])
    args = (e.ctypes.data_as(dp), len(e), ss.ctypes.data_as(ip), lay.ctypes.data_as(ip), ln.ctypes.data_as(dp),
            len(ss) - 1, lookup.ctypes.data_as(dp), len(lookup), s12, s13, dm)
    got, gen = np.empty((len(ss) - 1, len(e))), np.empty((len(ss) - 1, len(e)))
    n_ok = hm.hm_p1e_mirror(*args, 5.2, got.ctypes.data_as(dp))
    hm.hm_p1e(*args, gen.ctypes.data_as(dp))
    assert n_ok == int(np.sum(np.diff(ss) >= 3))          # every multi-layer chord of the default binning is mirrored
    mp = O.OscPars(dm, math.asin(math.sqrt(s12)), math.asin(math.sqrt(s13)))
    ref = O.osc_prob_earth(e, mp, lookup, tables["seg_start"], tables["seg_layer"], tables["seg_length"])
    assert np.max(np.abs(got - ref)) < 5e-12
    assert np.max(np.abs(got - gen)) < 5e-15


def test_mirror_eligibility_rejects_what_the_rotation_cannot_cover(hm, tables):
    """Not a palindrome, an even number of steps, or interior mirrored lengths too far apart for the second-order
    rotation: such paths take the generic walk (hm_p1e_mirror falls back per path and still matches the oracle)."""
    e = np.ascontiguousarray(tables["e_calc"])
    lookup = np.ascontiguousarray(tables["earth_lookup"])
    ss = np.array([0, 5, 10, 14, 19], dtype=np.int32)
    lay = np.array([0, 1, 2, 1, 0,   0, 1, 2, 1, 0,   0, 1, 1, 0,   0, 1, 2, 0, 1], dtype=np.int32)
    ln = np.array([300.0, 500.0, 2000.0, 500.0 + 1e-7, 300.002,      # eligible
                   300.0, 500.0, 2000.0, 500.1, 300.0,               # interior lengths 0.1 km apart
                   300.0, 500.0, 500.0, 300.0,                       # even number of steps
                   300.0, 500.0, 2000.0, 500.0, 300.0])              # layers not mirrored
    got = np.empty((4, len(e)))
    n_ok = hm.hm_p1e_mirror(e.ctypes.data_as(dp), len(e), ss.ctypes.data_as(ip), lay.ctypes.data_as(ip), ln.ctypes.data_as(dp), 4,
                            lookup.ctypes.data_as(dp), len(lookup), 0.307, 0.0224, 7.5e-5, 5.2, got.ctypes.data_as(dp))
    assert n_ok == 1
    # the oracle takes stored (detector -> surface) lists: applied layer m = stored layer n-1-m, applied length m = stored length m
    st_lay = np.concatenate([lay[a:b][::-1] for a, b in zip(ss[:-1], ss[1:])]) + 1
    mp = O.OscPars(7.5e-5, math.asin(math.sqrt(0.307)), math.asin(math.sqrt(0.0224)))
    ref = O.osc_prob_earth(e, mp, lookup, ss, st_lay, ln)
    assert np.max(np.abs(got - ref)) < 5e-12


def test_msw_c2m_matches_reference_formula(hm):
    mp = O.OscPars(6e-5, math.asin(math.sqrt(0.307)), math.asin(math.sqrt(0.0224)))
    for e in (1.1e-3, 8e-3, 2.09e-2):
        for ne in (20.0, 89.7, 101.0):
            _, c2m = O.msw_prob(e, mp, ne)
            got = hm.hm_msw_c2m(e, ne, 6e-5, 1 - 0.0224, 1 - 2 * 0.307)
            assert abs(got - c2m) < 1e-15
