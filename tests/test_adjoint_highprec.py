"""The adjoint sweep of csrc/solarosc_adjoint_math.cuh (the exact oscillation-stage derivatives behind
solarosc_eval_llh_grad) against 40-digit differentiation of the reference formula.

The reference differentiates total_llh with ForwardDiff (mcmc.jl:102-103): machine-precision derivatives of
osc.jl:360-388,436-477 and osc.jl:77-90.  Here the same formulas are written in mpmath (full complex Hamiltonian,
matrix exponentials) and differentiated numerically at 40 digits -- exact to ~1e-25 -- and the kernel's own step
functions, run on the host through tests/csrc/host_math_check.cu, must reproduce dP1e/d(sin2_th12, sin2_th13,
dm2_21, earth_norm_k) to 1e-10 relative.  A second test sweeps every path and energy against central differences
of the same host code (looser: the differences carry ~1e-7 of noise)."""
import ctypes as C
import math
import os
import shutil
import subprocess

import mpmath as mp
import numpy as np
import pytest

from oracle import oracle_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def hm():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    src = os.path.join(HERE, "csrc", "host_math_check.cu")
    out = os.path.join(HERE, "csrc", "_build", "libhostmath_adj.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(HERE, "..", "dune_solaroscfitter_b200", "csrc", f) for f in ("solarosc_device.cuh", "solarosc_adjoint_math.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-mfma", "-shared", "-o", out, src],
                       check=True, capture_output=True)
    L = C.CDLL(out)
    L.hm_adj_p1e_grad.argtypes = [dp, C.c_int, ip, ip, dp, C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp]
    L.hm_msw_deriv.argtypes = [C.c_double] * 5 + [dp, dp]
    return L


def _applied(t):
    ss = np.ascontiguousarray(t["seg_start"], dtype=np.int32)
    lay, ln = np.empty(ss[-1], dtype=np.int32), np.empty(ss[-1])
    for p in range(len(ss) - 1):
        s0, n = ss[p], ss[p + 1] - ss[p]
        for m in range(n):    # applied step m: layer index[n-1-m], length[m] (osc.jl:436-446, 455-456)
            lay[s0 + m] = t["seg_layer"][s0 + n - 1 - m] - 1
            ln[s0 + m] = t["seg_length"][s0 + m]
    return ss, lay, ln


def _adjoint(hm, t, s12, s13, dm, norms):
    ss, lay, ln = _applied(t)
    e = np.ascontiguousarray(t["e_calc"])
    lookup = np.ascontiguousarray(t["earth_lookup"], dtype=float)
    norms = np.ascontiguousarray(norms, dtype=float)
    n_paths, nl = len(ss) - 1, len(lookup)
    p1e, grad = np.empty((n_paths, len(e))), np.empty((n_paths, len(e), 3 + nl))
    hm.hm_adj_p1e_grad(e.ctypes.data_as(dp), len(e), ss.ctypes.data_as(ip), lay.ctypes.data_as(ip), ln.ctypes.data_as(dp), n_paths,
                       lookup.ctypes.data_as(dp), norms.ctypes.data_as(dp), nl, s12, s13, dm, p1e.ctypes.data_as(dp), grad.ctypes.data_as(dp))
    return p1e, grad


def mp_p1e(e, s12sq, s13sq, dm21, norms, lookup, layers, lengths):
    """P(nu_1 -> nu_e) of one path at one energy as a function of the sampled parameters: get_mixing_parameters
    (propagation_osc.jl:30-36), get_PMNS / get_H / osc_kernel / osc_reduce_from_lookup (osc.jl:57-71,360-388,436-477)."""
    f = mp.mpf
    th12, th13 = mp.asin(mp.sqrt(s12sq)), mp.asin(mp.sqrt(s13sq))
    th23, dcp, dm31 = mp.asin(mp.sqrt(f("0.5"))), f(-1.611), f("2.5e-3")
    c23, s23, c13, s13, c12, s12 = mp.cos(th23), mp.sin(th23), mp.cos(th13), mp.sin(th13), mp.cos(th12), mp.sin(th12)
    ep, em = mp.exp(1j * dcp), mp.exp(-1j * dcp)
    u = mp.matrix([[1, 0, 0], [0, c23, s23], [0, -s23, c23]]) * mp.matrix([[c13, 0, s13 * em], [0, 1, 0], [-s13 * ep, 0, c13]]) * \
        mp.matrix([[c12, s12, 0], [-s12, c12, 0], [0, 0, 1]])
    udag = u.transpose_conj()
    h_eff = u * mp.diag([0, dm21, dm31]) * udag
    x = mp.eye(3)
    idx, lens = layers[::-1], lengths[::-1]
    n = len(lens)
    for i in range(n):
        k = int(idx[n - 1 - i]) - 1
        rho_e = norms[k] * f(float(lookup[k])) * f(float(e))
        h = h_eff.copy()
        h[0, 0] += rho_e
        for d in range(3):
            h[d, d] -= rho_e
        x = x * mp.expm(-1j * f(O.PHASE_CONST) * (f(float(lens[i])) / f(float(e))) * h)
    a = (udag * x)[0, 0]
    return abs(a) ** 2


POINTS = [(0.307, 0.0224, 6e-5, [1.0] * 6),
          (0.2, 0.0255, 2.33e-5, [0.3, 1.9, 0.7, 1.2, 1.0, 1.0])]


@pytest.mark.parametrize("s12,s13,dm,norms", POINTS)
def test_adjoint_p1e_gradient_against_40_digit_differentiation(hm, tables, s12, s13, dm, norms):
    mp.mp.dps = 40
    p1e, grad = _adjoint(hm, tables, s12, s13, dm, norms)
    ss, sl, ln = tables["seg_start"], tables["seg_layer"], tables["seg_length"]
    e_all, lookup = np.asarray(tables["e_calc"]), np.asarray(tables["earth_lookup"])
    worst = 0.0
    for ipath, ie in ((0, 17), (4, 79), (23, 41), (49, 0)):          # 9, 7, 3 and 1 segments
        lay, lens = sl[ss[ipath]:ss[ipath + 1]], ln[ss[ipath]:ss[ipath + 1]]
        args = [mp.mpf(s12), mp.mpf(s13), mp.mpf(dm)] + [mp.mpf(v) for v in norms]
        crossed = sorted(set(int(k) - 1 for k in lay))
        dirs = [0, 1, 2] + [3 + crossed[0], 3 + crossed[-1]] + [3 + 5]   # two layers on the path and one that no default path reaches
        exact_p = mp_p1e(e_all[ie], args[0], args[1], args[2], args[3:], lookup, lay, lens)
        assert abs(p1e[ipath, ie] - float(exact_p)) < 2e-14
        for d in dirs:
            def f(v, d=d):
                a = list(args); a[d] = v
                return mp_p1e(e_all[ie], a[0], a[1], a[2], a[3:], lookup, lay, lens)
            exact = float(mp.diff(f, args[d], h=mp.mpf(10) ** -12 * max(abs(args[d]), mp.mpf(1e-6))))
            got = grad[ipath, ie, d]
            scale = max(abs(exact), 1e-3 * np.max(np.abs(grad[ipath, ie, :3])) if d < 3 else 1e-6)
            worst = max(worst, abs(got - exact) / scale)
            assert abs(got - exact) <= 1e-10 * scale, (ipath, ie, d, got, exact)
    print(f"worst relative deviation of the adjoint derivatives from 40-digit differentiation: {worst:.2e}")


def test_adjoint_gradient_against_central_differences_everywhere(hm, tables):
    """Every path and energy: central differences of the same host forward sweep (noise ~1e-7 relative to the largest
    derivative of the row)."""
    s12, s13, dm, norms = 0.33, 0.021, 7.4e-5, np.array([1.1, 0.8, 1.3, 0.9, 1.05, 1.0])
    _, grad = _adjoint(hm, tables, s12, s13, dm, norms)
    steps = [1e-6, 1e-6, 1e-10] + [1e-5] * 6
    for d, h in enumerate(steps):
        args_p, args_m = [s12, s13, dm, norms.copy()], [s12, s13, dm, norms.copy()]
        if d < 3:
            args_p[d] += h; args_m[d] -= h
        else:
            args_p[3][d - 3] += h; args_m[3][d - 3] -= h
        pp, _ = _adjoint(hm, tables, *args_p)
        pm, _ = _adjoint(hm, tables, *args_m)
        fd = (pp - pm) / (2 * h)
        scale = np.maximum(np.abs(fd), 1e-2 * np.max(np.abs(fd)) + 1e-12)
        assert np.max(np.abs(grad[:, :, d] - fd) / scale) < 2e-5, d
    assert np.all(grad[:, :, 3 + 5] == 0)          # no default path reaches the innermost layer


def test_adjoint_handles_the_clamped_mixing_angles(hm, tables):
    """get_mixing_parameters clamps sin^2 to [0, 1]: flat outside, so those derivatives vanish and the others are the
    ones of the clamped value."""
    _, g_out = _adjoint(hm, tables, 0.307, -0.1, 6e-5, np.ones(6))
    _, g_edge = _adjoint(hm, tables, 0.307, 0.0, 6e-5, np.ones(6))
    assert np.all(g_out[:, :, 1] == 0) and np.all(np.isfinite(g_out))
    assert np.array_equal(g_out, g_edge)
    _, g_hi = _adjoint(hm, tables, 1.2, 0.0224, 6e-5, np.ones(6))         # theta12 = pi/2: nu_1 decouples from nu_e
    assert np.all(g_hi == 0)


def test_msw_derivatives_against_40_digit_differentiation(hm):
    mp.mp.dps = 40

    def pday_k(e, ne, dm21, x, y):
        f = mp.mpf
        c2 = 1 - 2 * x
        beta = 2 * mp.sqrt(2) * f(ne) * f(O.G_F) * f(e) * (1 - y) / dm21          # osc.jl:77-90
        c2m = (c2 - beta) / mp.sqrt((c2 - beta) ** 2 + (1 - c2 ** 2))
        return (1 - y) ** 2 * (1 + c2 * c2m) / 2 + y ** 2, (1 - y) * c2m

    for e in (1.1e-3, 8e-3, 2.09e-2):
        for ne in (20.0, 89.7, 101.0):
            for (x, y, dm) in ((0.307, 0.0224, 6e-5), (0.21, 0.03, 1.4e-4)):
                v, d = np.empty(2), np.empty(6)
                hm.hm_msw_deriv(e, ne, dm, x, y, v.ctypes.data_as(dp), d.ctypes.data_as(dp))
                args = [mp.mpf(x), mp.mpf(y), mp.mpf(dm)]
                for which in range(2):
                    assert abs(v[which] - float(pday_k(e, ne, args[2], args[0], args[1])[which])) < 1e-15
                    for k in range(3):
                        def f(t, k=k, which=which):
                            a = list(args); a[k] = t
                            return pday_k(e, ne, a[2], a[0], a[1])[which]
                        exact = float(mp.diff(f, args[k], h=mp.mpf(10) ** -12 * args[k]))
                        assert abs(d[3 * which + k] - exact) <= 1e-12 * max(abs(exact), 1.0), (e, ne, which, k)
