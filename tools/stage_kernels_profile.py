#!/usr/bin/env python
"""One launch of every stage-separated kernel for an ncu capture (profiles/r01_stage_kernels_ncu.txt):
osc_kernel + rates_kernel (SOLAROSC_NO_FUSED=1), gen_kernel (inclusive + angular), the derived-mode reductions,
poisson_kernel.  Run under `ncu --set full -k regex:'osc_kernel|rates_kernel|gen_kernel|derived|poisson'`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SOLAROSC_NO_FUSED"] = "1"
from bench import load_tables, scan_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

t = load_tables()
data = oracle_np.asimov_data(t)
p, _ = scan_params(t, 32, float(t["truth"][1]))
p = np.ascontiguousarray(p[:16384])
with SolarOscEngine(t, data=data) as e:
    e.eval_llh(p)                                                   # osc_kernel + rates_kernel<llh>
    lref = {k: v[0] for k, v in e.eval_perbin(np.asarray(t["truth"], dtype=float)[None, :]).items()}
    e.eval_derived(p[:1024], np.ones(1024), lref)                   # rates_kernel<rates> + derived kernels
    rng = np.random.default_rng(0)
    m = rng.integers(10, 51, (64, 80400)).astype(float)
    e.poisson_llh(np.full_like(m, 30.0), m)                         # poisson_kernel
ext = dict(np.load(os.path.join(ROOT, "tests", "golden", "tables_ext.npz")))
tv = dict(t)
tv.update({k: ext[k] for k in ("n_cos", "resp_cc_incl", "eff_cc_incl", "resp_es_angular", "resp_bg_angular", "angular_mask")})
tv["eff_es_nue"], tv["eff_es_nuother"] = ext["eff_es_nue_boosted"], ext["eff_es_nuother_boosted"]
tv["inclusive"], tv["angular"] = np.int32(1), np.int32(1)
with SolarOscEngine(tv, data=oracle_np.asimov_data(tv)) as e:
    e.eval_llh(p)                                                   # osc_kernel + gen_kernel<llh>
print("done")
