"""Derived-mode quantities computed from batched rate arrays (host side, numpy).

The reference's `RunMode: derived` replays a stored chain one sample at a time
(`derive_variables_from_chain.jl:214-447`); with the engine the rates of a whole block of samples come
back from one `expected_rates_batch` call and the reductions below run over the batch axis.  Moving these
reductions onto the device is the next step; the formulas are the reference's.
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
