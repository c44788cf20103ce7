#!/usr/bin/env python
"""Throughput with fastFit: false (production-region integral over the 1000 solar-model radii, osc.jl:139-182)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_tables, scan_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

t = load_tables()
p, _ = scan_params(t, 64, float(t["truth"][1]))
for fast in (True, False):
    with SolarOscEngine(t, data=oracle_np.asimov_data(t, fast=fast), fast=fast) as e:
        out = np.empty(len(p))
        e.eval_llh(p, out=out)
        t0 = time.perf_counter()
        for _ in range(3):
            e.eval_llh(p, out=out)
        dt = (time.perf_counter() - t0) / 3
        print(f"fastFit = {fast}: {len(p) / dt:.3e} evaluations/s ({len(p)} points, {dt * 1e3:.1f} ms)")
