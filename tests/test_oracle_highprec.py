"""The oracle's double-precision arithmetic against the same formulas in 40-digit arithmetic (mpmath).

The reference cannot run here, so nothing pins the oracle's SEMANTICS beyond the reference's own known answers
(tests/test_oracle.py); what this file pins is its NUMERICS: the layered-Earth amplitude of osc.jl:360-388,436-477 --
full complex Hermitian Hamiltonian with theta23 and delta_CP, matrix exponentials, ordered 3x3 products -- evaluated
with matrix exponentials at 40 digits agrees with the oracle (LAPACK eigen-decompositions in binary64) to the
rounding noise that phases of up to 7e4 rad allow, and the kernels' real rank-one / state-vector algebra
(csrc/solarosc_device.cuh through the host harness) sits at the same distance from the exact value."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import oracle_np as O

POINTS = [(0.307, 0.0224, 6e-5, np.ones(6)),
          (0.2, 0.0255, 2.33e-5, np.array([0.3, 1.9, 0.7, 1.2, 1.0, 1.0])),
          (0.4, 0.03, 2e-4, np.full(6, 1.5))]
PATHS = [0, 4, 19, 23, 30, 49]          # 9, 7, 5, 3, 1, 1 segments
ENERGIES = [0, 17, 41, 79]


def mp_p1e(e, mpars, lookup, layers, lengths):
    """P(nu_1 -> nu_e) of one path at one energy, osc.jl:448-477 in 40-digit arithmetic."""
    mp.mp.dps = 40
    f = mp.mpf
    c23, s23 = mp.cos(f(mpars.th23)), mp.sin(f(mpars.th23))
    c13, s13 = mp.cos(f(mpars.th13)), mp.sin(f(mpars.th13))
    c12, s12 = mp.cos(f(mpars.th12)), mp.sin(f(mpars.th12))
    ep, em = mp.exp(1j * f(mpars.dcp)), mp.exp(-1j * f(mpars.dcp))
    u = mp.matrix([[1, 0, 0], [0, c23, s23], [0, -s23, c23]]) * mp.matrix([[c13, 0, s13 * em], [0, 1, 0], [-s13 * ep, 0, c13]]) * \
        mp.matrix([[c12, s12, 0], [-s12, c12, 0], [0, 0, 1]])
    udag = u.transpose_conj()
    h_eff = u * mp.diag([0, f(mpars.dm21), f(mpars.dm31)]) * udag
    x = mp.eye(3)
    idx, lens = layers[::-1], lengths[::-1]                      # reverse(...) at osc.jl:455-456
    n = len(lens)
    for i in range(n):
        rho_e = f(float(lookup[idx[n - 1 - i] - 1])) * f(float(e))
        h = h_eff.copy()
        h[0, 0] += rho_e
        for d in range(3):
            h[d, d] -= rho_e
        x = x * mp.expm(-1j * f(O.PHASE_CONST) * (f(float(lens[i])) / f(float(e))) * h)
    a = (udag * x)[0, 0]
    return float(abs(a) ** 2)


@pytest.mark.parametrize("s12,s13,dm,norms", POINTS)
def test_p1e_oracle_and_kernel_algebra_against_40_digits(tables, s12, s13, dm, norms):
    mpars = O.OscPars(dm, math.asin(math.sqrt(s12)), math.asin(math.sqrt(s13)))
    lookup = norms * tables["earth_lookup"]
    e_all = np.asarray(tables["e_calc"])
    ss, sl, ln = tables["seg_start"], tables["seg_layer"], tables["seg_length"]
    ref = O.osc_prob_earth(e_all[ENERGIES], mpars, lookup, ss, sl, ln)
    worst = 0.0
    for ip in PATHS:
        for k, ie in enumerate(ENERGIES):
            exact = mp_p1e(e_all[ie], mpars, lookup, sl[ss[ip]:ss[ip + 1]], ln[ss[ip]:ss[ip + 1]])
            worst = max(worst, abs(ref[ip, k] - exact))
            assert 0.0 <= exact <= 1.0
    # phases reach c L dm31 / E ~ 7e4 rad; one ulp of such a phase is 1.5e-11 rad
    assert worst < 2e-12


def test_kernel_algebra_is_as_close_to_exact_as_the_oracle(tables):
    import ctypes as C
    import os
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "csrc", "_build", "libhostmath.so")
    if not os.path.exists(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-mfma", "-shared", "-o", out,
                        os.path.join(here, "csrc", "host_math_check.cu")], check=True, capture_output=True)
    L = C.CDLL(out)
    dp, ip_ = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.hm_p1e.argtypes = [dp, C.c_int, ip_, ip_, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    s12, s13, dm, norms = POINTS[1]
    mpars = O.OscPars(dm, math.asin(math.sqrt(s12)), math.asin(math.sqrt(s13)))
    lookup = np.ascontiguousarray(norms * tables["earth_lookup"])
    ss = np.ascontiguousarray(tables["seg_start"], dtype=np.int32)
    lay, ln = np.empty(ss[-1], dtype=np.int32), np.empty(ss[-1])
    for pth in range(len(ss) - 1):     # applied step m: layer index[n-1-m], length[m] (osc.jl:436-446, 455-456)
        s0, n = ss[pth], ss[pth + 1] - ss[pth]
        for m in range(n):
            lay[s0 + m] = tables["seg_layer"][s0 + n - 1 - m] - 1
            ln[s0 + m] = tables["seg_length"][s0 + m]
    e = np.ascontiguousarray(tables["e_calc"])
    got = np.empty((len(ss) - 1, len(e)))
    L.hm_p1e(e.ctypes.data_as(dp), len(e), ss.ctypes.data_as(ip_), lay.ctypes.data_as(ip_), ln.ctypes.data_as(dp), len(ss) - 1,
             lookup.ctypes.data_as(dp), len(lookup), s12, s13, dm, got.ctypes.data_as(dp))
    ref = O.osc_prob_earth(e[ENERGIES], mpars, lookup, tables["seg_start"], tables["seg_layer"], tables["seg_length"])
    err_kernel = err_oracle = 0.0
    for ip in PATHS:
        for k, ie in enumerate(ENERGIES):
            exact = mp_p1e(e[ie], mpars, lookup, tables["seg_layer"][tables["seg_start"][ip]:tables["seg_start"][ip + 1]],
                           tables["seg_length"][tables["seg_start"][ip]:tables["seg_start"][ip + 1]])
            err_kernel = max(err_kernel, abs(got[ip, ie] - exact))
            err_oracle = max(err_oracle, abs(ref[ip, k] - exact))
    assert err_kernel < 2e-14 and err_oracle < 2e-12          # the kernel's Newton-polished eigen-system: ~1e-15
    print(f"max |P1e - exact|: kernel algebra {err_kernel:.2e}, oracle {err_oracle:.2e}")


def test_msw_formula_against_40_digits():
    mp.mp.dps = 40
    mpars = O.OscPars(6e-5, math.asin(math.sqrt(0.307)), math.asin(math.sqrt(0.0224)))
    for e in (1.1e-3, 8e-3, 2.09e-2):
        for ne in (20.0, 89.7, 101.0):
            p, c2m = O.msw_prob(e, mpars, ne)
            f = mp.mpf
            s2 = mp.sin(f(mpars.th13)) ** 2
            c2 = mp.cos(2 * f(mpars.th12))
            beta = 2 * mp.sqrt(2) * f(ne) * f(O.G_F) * f(e) * (1 - s2) / f(mpars.dm21)
            c2m_x = (c2 - beta) / mp.sqrt((c2 - beta) ** 2 + (1 - c2 ** 2))
            p_x = (1 - s2) ** 2 * (1 + c2 * c2m_x) / 2 + s2 ** 2
            assert abs(c2m - float(c2m_x)) < 1e-15 and abs(p - float(p_x)) < 1e-15
