#!/usr/bin/env python
"""Throughput of solarosc_eval_llh_multi: one host process, one handle per visible GPU, host buffers in and out."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_tables, scan_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine, eval_llh_multi  # noqa: E402
from oracle import oracle_np  # noqa: E402

t = load_tables()
data = oracle_np.asimov_data(t)
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p, _ = scan_params(t, grid, float(t["truth"][1]))
hp = torch.from_numpy(p).pin_memory().numpy()
n_dev = torch.cuda.device_count()
engines = [SolarOscEngine(t, data=data, device=d) for d in range(n_dev)]
for k in range(1, n_dev + 1):
    eval_llh_multi(engines[:k], hp)
    t0 = time.perf_counter()
    for _ in range(2):
        out = eval_llh_multi(engines[:k], hp)
    dt = (time.perf_counter() - t0) / 2
    print(f"{k} handle(s): {len(hp) / dt:.3e} evaluations/s from host buffers ({len(hp)} points, {dt * 1e3:.1f} ms)")
