#!/usr/bin/env python
"""Average per-CTA phase durations of the fused kernel (needs a library built with -DSOLAROSC_PHASE_TIMING).

usage: SOLAROSC_LIB=build_variants/lib_phase.so python tools/phase_timing.py [group]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from dune_solaroscfitter_b200 import _lib  # noqa: E402
from oracle import oracle_np  # noqa: E402
import bench  # noqa: E402

NAMES = ["prologue", "eigen+msw", "barrier(eigen)", "propagation (warp 0)", "barrier(propagation)", "fold vectors",
         "barrier(vectors)", "fold task + epilogue (warp 0)", "reduction barrier"]


def main():
    group = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    t = bench.load_tables()
    eng = SolarOscEngine(t, data=oracle_np.asimov_data(t), max_batch=32768 * group)
    params, n_cell = bench.scan_params(t, 40, float(t["truth"][1]))
    params = params[: 32768 * group]
    for _ in range(2):
        eng.eval_llh(params, group=group)
    L = _lib.load()
    L.solarosc_debug_read_phases.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    n = 16 * 4096
    buf = (C.c_longlong * n)()
    ns, nc = C.c_int(), C.c_int()
    L.solarosc_debug_read_phases(buf, C.byref(ns), C.byref(nc))
    a = np.frombuffer(buf, dtype=np.int64).reshape(nc.value, ns.value)
    a = a[a[:, 0] > 0]
    d = np.diff(a[:, :10], axis=1)
    tot = (a[:, 9] - a[:, 0]).mean()
    print(f"{len(a)} CTAs, mean cycles per CTA {tot:.0f}")
    pro = a[:, [0, 10, 11, 12, 1]]
    if np.all(pro[:, 1] > 0):
        dp = np.diff(pro, axis=1)
        for nm, col in zip(("  prologue: path tables to smem", "  prologue: zero padding", "  prologue: fold vectors (thread 0)", "  prologue: parameters + mixing"), dp.T):
            print(f"{nm:36s} {col.mean():9.0f}   (p10 {np.percentile(col, 10):.0f}, p90 {np.percentile(col, 90):.0f})")
    for i, nm in enumerate(NAMES):
        print(f"  {nm:32s} {d[:, i].mean():9.0f}  {100 * d[:, i].mean() / tot:5.1f}%   (p10 {np.percentile(d[:, i], 10):.0f}, p90 {np.percentile(d[:, i], 90):.0f})")


if __name__ == "__main__":
    main()
