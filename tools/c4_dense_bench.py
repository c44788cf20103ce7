"""C4: night-only CC likelihood on the dense 1000 x 1000 grid, 96 earth-norm draws per call (scratch tuning tool)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from dune_solaroscfitter_b200.engine import SolarOscEngine
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "dense1000"
td_ = dict(np.load(f"tests/golden/tables_{name}.npz"))
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 96
with SolarOscEngine(td_, device=0, es_mode=False, night_only=True) as ed:
    tr = np.asarray(td_["truth"], dtype=float)
    r0 = ed.eval_rates(tr[None, :])
    ed.set_data({k: r0[k][0] for k in ("ES_day", "ES_night", "CC_day", "CC_night")})
    pd_ = np.tile(tr, (nb, 1))
    pd_[:, 8:14] = np.clip(np.random.default_rng(3).normal(1.0, 0.12, (nb, 6)), 0.0, 2.0)
    pd_[:, 0], pd_[:, 2] = 0.4, 2e-4
    ed.set_timing(True)
    ed.eval_llh(pd_)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); l = ed.eval_llh(pd_); ts.append(time.perf_counter() - t0)
    tm = ed.last_timing() if hasattr(ed, "last_timing") else {}
    print("%s (SOLAROSC_OSC_CTAS_PER_SM=%s, NO_MIRROR=%s): %.0f eval/s" % (name, os.environ.get("SOLAROSC_OSC_CTAS_PER_SM", "default"), os.environ.get("SOLAROSC_NO_MIRROR", "0"), nb / min(ts)), "lnl[0:2]", l[:2], tm)
