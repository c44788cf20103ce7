#!/usr/bin/env python
"""C3 (HMC): value + gradient throughput of solarosc_eval_llh_grad from host buffers, B = 4096 chain points per call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from bench import load_tables  # noqa: E402
from conftest import draw_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

t = load_tables()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = draw_params(t, n, seed=5)
with SolarOscEngine(t, data=oracle_np.asimov_data(t)) as e:
    e.eval_llh_grad(p)
    t0 = time.perf_counter()
    for _ in range(reps):
        e.eval_llh_grad(p)
    dt = (time.perf_counter() - t0) / reps
    e.eval_llh(p)
    t0 = time.perf_counter()
    for _ in range(reps):
        e.eval_llh(p)
    dv = (time.perf_counter() - t0) / reps
print(f"B = {n}: value + gradient {dt * 1e3:.2f} ms per call = {n / dt:.3e} (lnL + gradient)/s; value only {dv * 1e3:.2f} ms = {n / dv:.3e}/s")
