"""Cycle accounting of pipe_kernel (diagnostic build -DSOLAROSC_PIPE_STATS): where the 16 warps of an SM spend their time.

  nvcc ... -DSOLAROSC_PIPE_STATS -o build_variants/libsolarosc_stats.so ...; SOLAROSC_LIB=build_variants/libsolarosc_stats.so python tools/pipe_stats.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_solaroscfitter_b200 import _lib  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import draw_params  # noqa: E402

t = dict(np.load(os.path.join(ROOT, "tests", "golden", "tables_escc.npz")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
group = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = draw_params(t, n)
L = _lib.load()
buf = (C.c_ulonglong * 16)()
with SolarOscEngine(t, data=oracle_np.asimov_data(t), device=0) as eng:
    f = (lambda: eng.eval_llh(p)) if group == 1 else (lambda: eng.eval_llh(p, group=group))
    have = hasattr(L, "solarosc_debug_read_pipe_stats")
    f()
    if have:
        L.solarosc_debug_read_pipe_stats(buf)
    t0 = time.perf_counter()
    f()
    dt = time.perf_counter() - t0
    if have:
        L.solarosc_debug_read_pipe_stats(buf)
if not have:
    print(f"{n} points, group {group}: {n / dt / 1e6:.2f} M/s (host-timed)")
    raise SystemExit(0)
v = np.array(list(buf), dtype=np.float64)
tot = v[7]
names = ["wait P<-E", "wait P<-F(tile)", "wait E<-P(table)", "wait F<-P", "run P", "run E", "run F", "warp lifetime"]
print(f"{n} points, group {group}: {n / dt / 1e6:.2f} M/s (host-timed, stats build)")
for k, nm in enumerate(names):
    print(f"  {nm:18s} {v[k]:.3e} cycles  {100 * v[k] / tot:5.1f} % of warp lifetime")
print(f"  unaccounted        {100 * (tot - v[:7].sum()) / tot:5.1f} %")
