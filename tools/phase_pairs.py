#!/usr/bin/env python
"""Where in its own period is the co-resident CTA when a CTA of the fused kernel starts?  (needs a library built with
-DSOLAROSC_PHASE_TIMING; slot 15 of the phase record is the SM id, slot 0 / 9 the CTA's first / last clock64 on that SM.)

Two CTAs share an SM.  If their phases were uncorrelated the offsets would be uniform on [0, 1); a peak near 0 means they
run in lock step (eigen table against eigen table, propagation against propagation).

usage: SOLAROSC_LIB=build_variants/phase.so python tools/phase_pairs.py
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


def main():
    t = bench.load_tables()
    eng = SolarOscEngine(t, data=oracle_np.asimov_data(t), max_batch=32768)
    params, _ = bench.scan_params(t, 40, float(t["truth"][1]))
    params = params[:32768]
    for _ in range(2):
        eng.eval_llh(params)
    L = _lib.load()
    L.solarosc_debug_read_phases.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    buf = (C.c_longlong * (16 * 4096))()
    ns, nc = C.c_int(), C.c_int()
    L.solarosc_debug_read_phases(buf, C.byref(ns), C.byref(nc))
    a = np.frombuffer(buf, dtype=np.int64).reshape(nc.value, ns.value)
    a = a[a[:, 0] > 0]
    start, end, sm = a[:, 0], a[:, 9], a[:, 15]
    offs, durs = [], []
    for s in np.unique(sm):
        k = np.where(sm == s)[0]
        k = k[np.argsort(start[k])]
        for i in k:
            # the CTA of this SM that was running when CTA i started
            other = [j for j in k if j != i and start[j] < start[i] < end[j]]
            if len(other) == 1:
                j = other[0]
                offs.append((start[i] - start[j]) / (end[j] - start[j]))
                durs.append(end[i] - start[i])
    offs = np.array(offs)
    h, _ = np.histogram(offs, bins=10, range=(0, 1))
    print(f"{len(a)} CTAs on {len(np.unique(sm))} SMs, {len(offs)} starts with exactly one co-resident CTA; mean CTA duration {np.mean(end - start):.0f} cycles")
    print("offset of a starting CTA within its neighbour's period (tenths):", (h / max(1, len(offs))).round(3).tolist())
    durs = np.array(durs)
    for lo in np.arange(0, 1, 0.2):
        m = (offs >= lo) & (offs < lo + 0.2)
        if m.any():
            print(f"  starts at {lo:.1f}-{lo + 0.2:.1f} of the neighbour's period: {m.sum():5d} CTAs, mean own duration {durs[m].mean():.0f} cycles")


if __name__ == "__main__":
    main()
