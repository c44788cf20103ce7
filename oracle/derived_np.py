"""TEST INFRASTRUCTURE (checker, not product): derived-mode quantities computed from batched rate arrays (numpy).
Only tests/ import this; the product computes the same sums on the device (`solarosc_eval_derived`).

The reference's `RunMode: derived` replays a stored chain one sample at a time
(`derive_variables_from_chain.jl:214-447`); with the engine the rates of a whole block of samples come
back from one `expected_rates_batch` call and the reductions below run over the batch axis.  The engine does
the same reductions on the device (`SolarOscEngine.eval_derived` -> `solarosc_eval_derived`); the numpy versions
here are the host-side statement of the formulas and what the tests compare the device sums with.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def day_night_asymmetry(day: np.ndarray, night: np.ndarray, bg_tot: np.ndarray, index: int) -> np.ndarray:
    """2 (D - N) / (D + N) from background-subtracted sums above threshold, per sample.

    `day` [B, n_ereco], `night` [B, n_cosz, n_ereco], `bg_tot` [B, n_ereco]; `index` is the 1-based first
    bin (derive_variables_from_chain.jl:236-258: the background is split half day / half night, and a
    zero denominator gives 0)."""
    k = index - 1
    bg = bg_tot[:, k:].sum(axis=1)
    n_tot = night[:, :, k:].sum(axis=(1, 2)) - 0.5 * bg
    d_tot = day[:, k:].sum(axis=1) - 0.5 * bg
    den = d_tot + n_tot
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den == 0, 0.0, 2 * (d_tot - n_tot) / np.where(den == 0, 1.0, den))


def weighted_moments(rates: np.ndarray, weights: np.ndarray, index: int) -> Tuple[np.ndarray, np.ndarray]:
    """sum_i w_i r_i and sum_i w_i r_i^2 over the sample axis, bins below threshold left at zero
    (derive_variables_from_chain.jl:285-330)."""
    k = index - 1
    w = np.asarray(weights, dtype=np.float64).reshape((-1,) + (1,) * (rates.ndim - 1))
    m1 = np.zeros(rates.shape[1:])
    m2 = np.zeros(rates.shape[1:])
    m1[..., k:] = (w * rates[..., k:]).sum(axis=0)
    m2[..., k:] = (w * rates[..., k:] ** 2).sum(axis=0)
    return m1, m2


def replay_block_sums(rates: dict, perbin: dict, lref: dict, params: np.ndarray, weights: np.ndarray,
                      index_es: int, index_cc: int) -> dict:
    """The replay loop of derive_variables_from_chain.jl:260-444 over a block of samples, written as that loop.

    `rates` / `perbin` hold the per-sample arrays (ES_day [B, n], ES_night [B, n_cosz, n], ... and BG_*_tot for the
    rates), `lref` the per-bin terms the loop centres on.  Returns the same keys as SolarOscEngine.eval_derived."""
    B = len(weights)
    nz = rates["ES_night"].shape[1]
    out = {"scalars": np.zeros(21), "ltot": np.zeros((B, 4))}
    for ch, idx in (("ES", index_es), ("CC", index_cc)):
        n = rates[ch + "_day"].shape[1]
        out["mom_" + ch] = np.zeros((6, 1 + nz, n))
        out["bg_mom_" + ch] = np.zeros((2, n))
    for i in range(B):
        w, t1, t2 = float(weights[i]), float(params[i, 0]), float(params[i, 2])
        sc = out["scalars"]
        sc[0] += w; sc[1] += w * t1; sc[2] += w * t1 ** 2; sc[3] += w * t2; sc[4] += w * t2 ** 2
        for c, (ch, idx) in enumerate((("ES", index_es), ("CC", index_cc))):
            k = idx - 1
            r = np.vstack([rates[ch + "_day"][i][None, :], rates[ch + "_night"][i]])
            l = np.vstack([perbin[ch + "_day"][i][None, :], perbin[ch + "_night"][i]]) - \
                np.vstack([np.asarray(lref[ch + "_day"])[None, :], np.asarray(lref[ch + "_night"])])
            m = out["mom_" + ch]
            m[0][:, k:] += w * r[:, k:]
            m[1][:, k:] += w * r[:, k:] ** 2
            m[2] += w * l; m[3] += w * l ** 2; m[4] += w * l * t1; m[5] += w * l * t2
            bg = 0.5 * rates["BG_" + ch + "_tot"][i]
            out["bg_mom_" + ch][0] += w * bg
            out["bg_mom_" + ch][1] += w * bg ** 2
            for g, tot in ((2 * c, l[1:].sum()), (2 * c + 1, l[0].sum())):     # night, day
                out["ltot"][i, g] = tot
                sc[5 + 4 * g: 9 + 4 * g] += np.array([w * tot, w * tot ** 2, w * tot * t1, w * tot * t2])
    return out


def replay_block_sums_angular(rates: dict, perbin: dict, lref: dict, params: np.ndarray, weights: np.ndarray, index_es: int) -> dict:
    """The ES part of the same loop with the angular reconstruction (derive_variables_from_chain.jl:239-258, 288-296,
    315-318, 366-433): ES_day [B, n_cos, n], ES_night [B, n_cos, n, n_cosz] per sample.  Returns mom_ES_day
    [6, n_cos, n], mom_ES_night [6, n_cos, n, n_cosz], ES_day [B, n] (the day spectrum summed over the cos-scatter bins)
    and the ES columns of ltot [B, 2] (night, day)."""
    B = len(weights)
    k = index_es - 1
    md, mn = np.zeros((6,) + rates["ES_day"].shape[1:]), np.zeros((6,) + rates["ES_night"].shape[1:])
    es_day, ltot = np.zeros((B, rates["ES_day"].shape[2])), np.zeros((B, 2))
    for i in range(B):
        w, t1, t2 = float(weights[i]), float(params[i, 0]), float(params[i, 2])
        rd, rn = rates["ES_day"][i], rates["ES_night"][i]
        md[0][:, k:] += w * rd[:, k:]; md[1][:, k:] += w * rd[:, k:] ** 2
        mn[0][:, k:, :] += w * rn[:, k:, :]; mn[1][:, k:, :] += w * rn[:, k:, :] ** 2
        ld = perbin["ES_day"][i] - np.asarray(lref["ES_day"])
        ln = perbin["ES_night"][i] - np.asarray(lref["ES_night"])
        for m, l in ((md, ld), (mn, ln)):
            m[2] += w * l; m[3] += w * l ** 2; m[4] += w * l * t1; m[5] += w * l * t2
        es_day[i] = rd.sum(axis=0)
        ltot[i] = ln.sum(), ld.sum()
    return {"mom_ES_day": md, "mom_ES_night": mn, "ES_day": es_day, "ltot_ES": ltot}
