#!/usr/bin/env python
"""Per-call latency of the host-buffer entry points for MCMC-sized batches (C1 / C3: B = 1, 10, 100, 4096)."""
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
with SolarOscEngine(t, data=oracle_np.asimov_data(t)) as e:
    for B in (1, 10, 100, 4096):
        p = draw_params(t, B, seed=B)
        out = np.empty(B)
        for fn, name, reps in ((lambda: e.eval_llh(p, out=out), "eval_llh", 300), (lambda: e.eval_llh_grad(p), "eval_llh_grad", 50)):
            for _ in range(5):
                fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            dt = (time.perf_counter() - t0) / reps
            print(f"B = {B:5d}  {name:14s} {dt * 1e6:9.1f} us per call  {B / dt:12.0f} evaluations/s")
