"""Minimax fits behind sincos_qt (csrc/solarosc_device.cuh): for f in [-1/2, 1/2] quarter turns,
    sin(pi/2 f) = f * Ps(f^2),  cos(pi/2 f) = Pc(f^2),  Pc(0) = 1 exactly.
Remez exchange in 60-digit arithmetic; prints the double coefficients and the achieved errors (evaluated
in double with FMAs as the kernel does).  Usage: python tools/fit_quarter_turn_trig.py [deg_sin deg_cos]"""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60
HALF_PI = mp.pi / 2
ZMAX = mp.mpf(1) / 4 * (1 + mp.mpf(2) ** -20)   # |f| may exceed 1/2 by rounding of the magic-number trick


def g_sin(z):
    if z == 0:
        return HALF_PI
    r = mp.sqrt(z)
    return mp.sin(HALF_PI * r) / r


def g_cos(z):   # (cos - 1)/z so that Pc(0) = 1 is exact
    if z == 0:
        return -HALF_PI ** 2 / 2
    return (mp.cos(HALF_PI * mp.sqrt(z)) - 1) / z


def remez(g, deg, weight=lambda z: mp.mpf(1), iters=12):
    n = deg + 2
    xs = [ZMAX * (1 - mp.cos(mp.pi * k / (n - 1))) / 2 for k in range(n)]
    for _ in range(iters):
        A = mp.matrix(n, n)
        b = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(deg + 1):
                A[i, j] = x ** j
            A[i, deg + 1] = (-1) ** i / weight(x)
            b[i] = g(x)
        sol = mp.lu_solve(A, b)
        coef = [sol[j] for j in range(deg + 1)]
        err = lambda x: (mp.polyval(coef[::-1], x) - g(x)) * weight(x)
        # new extrema: dense scan + local refinement between sign changes
        grid = [ZMAX * k / 4000 for k in range(4001)]
        vals = [err(x) for x in grid]
        ext = []
        for k in range(len(grid)):
            l = vals[k - 1] if k > 0 else None
            r = vals[k + 1] if k + 1 < len(grid) else None
            v = vals[k]
            if (l is None or abs(v) >= abs(l)) and (r is None or abs(v) >= abs(r)):
                if ext and mp.sign(vals[ext[-1]]) == mp.sign(v):
                    if abs(v) > abs(vals[ext[-1]]):
                        ext[-1] = k
                else:
                    ext.append(k)
        if len(ext) < n:
            break
        # keep the n largest consecutive alternating extrema
        while len(ext) > n:
            if abs(vals[ext[0]]) < abs(vals[ext[-1]]):
                ext.pop(0)
            else:
                ext.pop()
        xs = [grid[k] for k in ext]
    return coef, max(abs(v) for v in vals)


def main():
    ds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dc = int(sys.argv[2]) if len(sys.argv) > 2 else 5   # degree of (cos-1)/z: cos has dc+1 coefficients besides the 1
    cs, es = remez(g_sin, ds)
    cc, ec = remez(g_cos, dc)
    print("sin: deg %d in z, minimax abs error of Ps %.3e" % (ds, float(es)))
    print("cos: deg %d in z for (cos-1)/z, minimax abs error %.3e (x z <= 1/4)" % (dc, float(ec)))
    S = [float(c) for c in cs]
    Cc = [float(c) for c in cc]
    print("S =", ", ".join("%.20e" % c for c in S))
    print("C =", ", ".join("%.20e" % c for c in Cc))
    # double evaluation as in the kernel (numpy has no fma; emulate with longdouble to keep rounding honest)
    f = np.linspace(-0.5, 0.5, 200001)
    z = f * f
    ps = np.full_like(z, S[-1])
    for c in S[-2::-1]:
        ps = ps * z + c
    pc = np.full_like(z, Cc[-1])
    for c in Cc[-2::-1]:
        pc = pc * z + c
    s = f * ps
    c = pc * z + 1.0
    print("double eval: max |sin err| %.3e  max |cos err| %.3e" % (np.max(np.abs(s - np.sin(np.pi / 2 * f))), np.max(np.abs(c - np.cos(np.pi / 2 * f)))))


main()
