#!/usr/bin/env python
"""C5: derived-mode replay of a stored 1.5e5-sample chain (derive_variables_from_chain.jl): asymmetries for every sample,
block sums for the thinned-in ones, from host buffers."""
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
n = 150_000
p = draw_params(t, n, seed=5)
w = np.ones(n)
with SolarOscEngine(t, data=oracle_np.asimov_data(t)) as e:
    lref = {k: v[0] for k, v in e.eval_perbin(np.asarray(t["truth"], dtype=float)[None, :]).items()}
    e.eval_asymmetry(p[:4096]); e.eval_derived(p[:4096], w[:4096], lref)
    t0 = time.perf_counter(); e.eval_asymmetry(p); ta = time.perf_counter() - t0
    t0 = time.perf_counter(); e.eval_derived(p, w, lref); td = time.perf_counter() - t0
    t0 = time.perf_counter(); e.eval_rates(p[:8192]); tr = (time.perf_counter() - t0) * n / 8192
print(f"{n} chain samples: asymmetries {ta * 1e3:.0f} ms ({n / ta:.3e}/s), block sums {td * 1e3:.0f} ms ({n / td:.3e}/s); "
      f"rate arrays to the host instead would take {tr * 1e3:.0f} ms ({n / tr:.3e}/s, 49 kB per sample)")
