#!/usr/bin/env python
"""One launch of the fused kernels other than llh_kernel for an ncu capture (profiles/r02_other_kernels_ncu.txt):
grad_kernel + adjoint_kernel (value + gradient, exclusive analysis), gen_llh_kernel and gen_grad_kernel (inclusive + angular).
Run under `ncu --set full -k regex:'grad_kernel|adjoint_kernel|gen_llh_kernel' -c 6`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import load_tables, scan_params  # noqa: E402
from conftest import draw_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

t = load_tables()
p4096 = draw_params(t, 4096, seed=5)
with SolarOscEngine(t, data=oracle_np.asimov_data(t)) as e:
    e.eval_llh_grad(p4096)                                          # osc_kernel + grad_kernel + adjoint_kernel
ext = dict(np.load(os.path.join(ROOT, "tests", "golden", "tables_ext.npz")))
tv = dict(t)
tv.update({k: ext[k] for k in ("n_cos", "resp_cc_incl", "eff_cc_incl", "resp_es_angular", "resp_bg_angular", "angular_mask")})
tv["eff_es_nue"], tv["eff_es_nuother"] = ext["eff_es_nue_boosted"], ext["eff_es_nuother_boosted"]
tv["inclusive"], tv["angular"] = np.int32(1), np.int32(1)
p, _ = scan_params(t, 35, float(t["truth"][1]))
with SolarOscEngine(tv, data=oracle_np.asimov_data(tv)) as e:
    e.eval_llh(np.ascontiguousarray(p[:32768]))                     # gen_llh_kernel
    e.eval_llh_grad(p4096)                                          # osc_kernel + gen_grad_kernel + adjoint_kernel
print("done")
