"""CPU oracle (numpy twin) for the ES+CC day/night likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement of the reference's
algorithm, used as the checker by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
path (``dune_solaroscfitter_b200``) never imports it.

PARITY PINNING.  The reference is Julia and cannot run in this image, and its own
tests pin only the Poisson term (``tests/test_likelihood_correctness.jl:5-17``)
and the earth-norm bound (``:164-182``); those known answers are checked in
``tests/test_oracle.py``.  No reference test or stored fixture pins the MSW
probability, the Earth-matter P1e, the response fold or the background
normalisation, so for those stages this oracle is "parity unpinned": it is
validated by construction (each function follows the cited lines literally),
by agreement with the independent C restatement in ``oracle/solarosc_oracle.c``
and by physics known answers (vacuum limit, unitarity).

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Evaluation order follows the Julia expressions where the
language defines one; BLAS-backed products (``R' * osc``) have no defined
order in the reference either.

Parameter vector layout (same as the C ABI, include/solarosc_b200.h):
  [sin2_th12, sin2_th13, dm2_21, integrated_8B_flux, integrated_HEP_flux,
   cc_xsec_norm, cc_xsec_tilt, cc_xsec_curv, earth_norm_1..L,
   ES_bg_norm_1..nES, CC_bg_norm_1..nCC]
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple, Optional

import numpy as np

G_F = 5.4489e-5                      # osc.jl:51
PHASE_CONST = 2.5338653580781976     # osc.jl:361
DM2_31, THETA_23, DELTA_CP = 2.5e-3, math.asin(math.sqrt(0.5)), -1.611   # osc.jl:44


class OscPars(NamedTuple):           # osc.jl:33-47
    dm21: float
    th12: float
    th13: float
    dm31: float = DM2_31
    th23: float = THETA_23
    dcp: float = DELTA_CP


class ParamView:
    """Named access into the flat parameter vector."""

    def __init__(self, p, n_layers: int, n_es: int, n_cc: int):
        p = np.asarray(p, dtype=np.float64)
        self.sin2_th12, self.sin2_th13, self.dm2_21 = p[0], p[1], p[2]
        self.integrated_8B_flux, self.integrated_HEP_flux = p[3], p[4]
        self.cc_xsec_norm, self.cc_xsec_tilt, self.cc_xsec_curv = p[5], p[6], p[7]
        self.earth_norms = p[8:8 + n_layers]
        self.es_bg_norms = p[8 + n_layers:8 + n_layers + n_es]
        self.cc_bg_norms = p[8 + n_layers + n_es:8 + n_layers + n_es + n_cc]


def n_params(t) -> int:
    return 8 + len(t["earth_lookup"]) + int(np.count_nonzero(t["bg_es_par"])) + int(np.count_nonzero(t["bg_cc_par"]))


def view(p, t) -> ParamView:
    return ParamView(p, len(t["earth_lookup"]), int(np.count_nonzero(t["bg_es_par"])),
                     int(np.count_nonzero(t["bg_cc_par"])))


# ---------------------------------------------------------------------------
# oscillation physics  (src/oscillations/osc.jl)
# ---------------------------------------------------------------------------
def get_mixing_parameters(pv: ParamView) -> OscPars:
    """propagation_osc.jl:30-36."""
    return OscPars(float(pv.dm2_21),
                   math.asin(math.sqrt(min(max(pv.sin2_th12, 0.0), 1.0))),
                   math.asin(math.sqrt(min(max(pv.sin2_th13, 0.0), 1.0))))


def get_pmns(mp: OscPars) -> np.ndarray:
    """U = U23 * U13(delta) * U12 with the column-major literal placement of osc.jl:57-64."""
    c23, s23 = math.cos(mp.th23), math.sin(mp.th23)
    c13, s13 = math.cos(mp.th13), math.sin(mp.th13)
    c12, s12 = math.cos(mp.th12), math.sin(mp.th12)
    ep, em = np.exp(1j * mp.dcp), np.exp(-1j * mp.dcp)
    u1 = np.array([[1, 0, 0], [0, c23, s23], [0, -s23, c23]], dtype=complex)
    u2 = np.array([[c13, 0, s13 * em], [0, 1, 0], [-s13 * ep, 0, c13]], dtype=complex)
    u3 = np.array([[c12, s12, 0], [-s12, c12, 0], [0, 0, 1]], dtype=complex)
    return u1 @ u2 @ u3


def msw_prob(energy, mp: OscPars, n_e):
    """Adiabatic MSW survival probability and cos(2 theta12_matter) (osc.jl:77-90)."""
    s2th13 = math.sin(mp.th13) ** 2
    c2th12 = math.cos(2 * mp.th12)
    acc = 2 * math.sqrt(2) * n_e * G_F * energy * (1 - s2th13)
    beta = acc / mp.dm21
    c2m = (c2th12 - beta) / np.sqrt((c2th12 - beta) ** 2 + (1 - c2th12 ** 2))
    s13m = math.sqrt(s2th13)
    probs = 1 / 2 * (1 - s2th13) * (1 - s13m ** 2) * (1 + c2th12 * c2m) + s2th13 * s13m ** 2
    return probs, c2m


def get_h(h_vac: np.ndarray, e: float, rho: float):
    """Matter Hamiltonian eigen-system (osc.jl:370-388, neutrino branch)."""
    h = np.array(h_vac, dtype=complex)
    h[0, 0] += rho * e
    for i in range(3):
        h[i, i] -= rho * e
    vals, vecs = np.linalg.eigh(h, UPLO="U")   # Hermitian() reads the upper triangle
    return vecs, vals


def osc_kernel(v: np.ndarray, lam: np.ndarray, e: float, l: float) -> np.ndarray:
    """One constant-density step (osc.jl:360-364); note the adjoint on the phase row."""
    phase = np.exp((PHASE_CONST * 1j * (l / e)) * lam)
    tmp = v * np.conj(phase)[None, :]
    return tmp @ v.conj().T


def osc_prob_earth(e_calc, mp: OscPars, lookup, seg_start, seg_layer, seg_length) -> np.ndarray:
    """P(nu_1 -> nu_e) per (path, energy) through the layered Earth (osc.jl:448-477, 436-446)."""
    u = get_pmns(mp)
    udag = u.conj().T
    h_eff = u @ np.diag(np.array([0.0, mp.dm21, mp.dm31], dtype=complex)) @ udag
    n_paths = len(seg_start) - 1
    p1e = np.empty((n_paths, len(e_calc)))
    for ie, e in enumerate(e_calc):
        buf = [get_h(h_eff, float(e), float(rho)) for rho in lookup]
        for ip in range(n_paths):
            idx = seg_layer[seg_start[ip]:seg_start[ip + 1]][::-1]      # reverse(...) at :455
            lens = seg_length[seg_start[ip]:seg_start[ip + 1]][::-1]    # reverse(...) at :456
            n = len(lens)
            v, lam = buf[idx[n - 1] - 1]
            x = osc_kernel(v, lam, float(e), float(lens[0]))
            for i in range(1, n):
                v, lam = buf[idx[n - 1 - i] - 1]
                x = x @ osc_kernel(v, lam, float(e), float(lens[i]))
            p1e[ip, ie] = abs((udag @ x)[0, 0]) ** 2
    return p1e


def osc_prob_both_fast(e_calc, p1e, mp: OscPars, avg_ne: float):
    """Day probability at the mean production density + IYSW night formula (osc.jl:185-207)."""
    prob_day, c2m = msw_prob(np.asarray(e_calc), mp, avg_ne)
    cos2_13 = math.cos(mp.th13) ** 2
    p0 = math.cos(mp.th12) ** 2 * cos2_13
    prob_night = prob_day[None, :] + cos2_13 * c2m[None, :] * (p1e - p0)
    return prob_day, prob_night


def osc_prob_both_slow(e_calc, p1e, mp: OscPars, n_e, prod):
    """Production-region integral radius by radius (osc.jl:139-182)."""
    e_calc = np.asarray(e_calc)
    prob_day, c2m = msw_prob(e_calc[:, None], mp, np.asarray(n_e)[None, :])
    cos2_13 = math.cos(mp.th13) ** 2
    p0 = math.cos(mp.th12) ** 2 * cos2_13
    night = np.zeros_like(p1e)
    for i in range(len(n_e)):
        dp = cos2_13 * c2m[:, i][None, :] * (p1e - p0)
        night += prod[i] * (prob_day[:, i][None, :] + dp)
    total = np.sum(prod)
    return (prob_day @ prod) / total, night / total


def block_average(a: np.ndarray, block) -> np.ndarray:
    """Mean over consecutive blocks (propagation_core.jl:8-56)."""
    if a.ndim == 1:
        bs = int(block)
        if len(a) % bs:
            raise ValueError("Vector length must be a multiple of block size")
        return a.reshape(-1, bs).sum(axis=1) / bs
    br, bc = block
    nr, nc = a.shape
    if nr % br or nc % bc:
        raise ValueError("Matrix dimensions must be multiples of block dimensions")
    return a.reshape(nr // br, br, nc // bc, bc).sum(axis=(1, 3)) / (br * bc)


def compute_shared_osc_probs(p, t, fast: bool = True, earth_uncertainty: bool = True) -> Dict[str, np.ndarray]:
    """propagation_main.jl:47-54 -> propagation_osc.jl:52-74, 94-156."""
    pv = view(p, t)
    mp = get_mixing_parameters(pv)
    lookup = pv.earth_norms * t["earth_lookup"] if earth_uncertainty else t["earth_lookup"]
    p1e = osc_prob_earth(t["e_calc"], mp, lookup, t["seg_start"], t["seg_layer"], t["seg_length"])
    if fast:
        d8, n8 = osc_prob_both_fast(t["e_calc"], p1e, mp, float(t["avg_ne_8b"]))
        dh, nh = osc_prob_both_fast(t["e_calc"], p1e, mp, float(t["avg_ne_hep"]))
    else:
        d8, n8 = osc_prob_both_slow(t["e_calc"], p1e, mp, t["solar_ne"], t["solar_prod_8b"])
        dh, nh = osc_prob_both_slow(t["e_calc"], p1e, mp, t["solar_ne"], t["solar_prod_hep"])
    ns = int(t["n_cosz_sub"])
    out = dict(nue_8B_day=block_average(d8, 2), nue_hep_day=block_average(dh, 2),
               nue_8B_night=block_average(n8, (ns, 2)), nue_hep_night=block_average(nh, (ns, 2)))
    for k in ("8B_day", "hep_day", "8B_night", "hep_night"):
        out["nuother_" + k] = 1 - out["nue_" + k]
    out["p1e"] = p1e
    return out


# ---------------------------------------------------------------------------
# event rates  (src/propagation/*.jl)
# ---------------------------------------------------------------------------
def compute_oscillated_samples(t, pv: ParamView, osc):
    """Flux x probability x exposure (x CC cross-section shape) (propagation_osc.jl:159-210)."""
    f8, fh = pv.integrated_8B_flux, pv.integrated_HEP_flux
    wz = t["exposure_weights"][:, None]

    def day(u8, uh, p8, ph):
        return u8 * p8 * f8 + uh * ph * fh

    def night(u8, uh, p8, ph):
        return u8[None, :] * (f8 * p8 * wz) + uh[None, :] * (fh * ph * wz)

    es = dict(
        nue_day=day(t["unosc_es_nue_8b"], t["unosc_es_nue_hep"], osc["nue_8B_day"], osc["nue_hep_day"]),
        nue_night=night(t["unosc_es_nue_8b"], t["unosc_es_nue_hep"], osc["nue_8B_night"], osc["nue_hep_night"]),
        nuother_day=day(t["unosc_es_nuother_8b"], t["unosc_es_nuother_hep"], osc["nuother_8B_day"], osc["nuother_hep_day"]),
        nuother_night=night(t["unosc_es_nuother_8b"], t["unosc_es_nuother_hep"], osc["nuother_8B_night"], osc["nuother_hep_night"]),
    )
    ln = t["log_e_norm"]
    shape = np.exp(pv.cc_xsec_tilt * ln + pv.cc_xsec_curv * ln ** 2) * pv.cc_xsec_norm
    c8, ch = t["unosc_cc_8b"] * shape, t["unosc_cc_hep"] * shape
    cc = dict(day=day(c8, ch, osc["nue_8B_day"], osc["nue_hep_day"]),
              night=night(c8, ch, osc["nue_8B_night"], osc["nue_hep_night"]))
    return es, cc


def apply_day_response(osc_vec, resp, eff, scale=0.5):
    """propagation_core.jl:73-75."""
    return scale * ((resp.T @ osc_vec) * eff)


def apply_night_response(osc_rows, resp, eff, scale=0.5):
    """propagation_core.jl:77-79 (row by row in the reference)."""
    return np.vstack([scale * ((row @ resp) * eff) for row in osc_rows])


def es_conversion(side: int, norm: float) -> float:
    """propagation_bg.jl:6-18."""
    if side == 0:
        area = 2 * 12e2 * 64e2
    elif side == 1:
        area = 2 * 12e2 * 12e2
    else:
        area = 2 * 12e2 * 64e2 + 2 * 12e2 * 12e2
    return area * norm


def normalize_backgrounds(t, pv: ParamView):
    """propagation_bg.jl:36-83 (every ES component uses ES_bg_norm_1, :56-57)."""
    bg_es = np.zeros(t["raw_bg_es"].shape[1])
    for i, beh in enumerate(t["bg_es_par"]):
        if beh != 0:
            bg_es = bg_es + t["raw_bg_es"][i] * es_conversion(int(t["bg_es_sides"][i]), pv.es_bg_norms[0])
        else:
            bg_es = bg_es + t["raw_bg_es"][i]
    bg_cc = np.zeros(t["raw_bg_cc"].shape[1])
    k = 0
    for i, beh in enumerate(t["bg_cc_par"]):
        if beh != 0:
            bg_cc = bg_cc + t["raw_bg_cc"][i] * (pv.cc_bg_norms[k] / 2.2e-6)
            k += 1
        else:
            bg_cc = bg_cc + t["raw_bg_cc"][i]
    return bg_es, bg_cc


def _flag(t, key) -> int:
    return int(t[key]) if key in t else 0


def propagate_samples(p, t, fast=True, earth_uncertainty=True, precomputed_osc=None, es_mode=True, cc_mode=True):
    """propagateSamples (propagation_main.jl:56-157), all branches.  The optional table keys `inclusive`
    (0 exclusive, 1 inclusive, 2 semi-inclusive) and `angular` select them; with `angular` the ES arrays are
    (n_cos, n_ereco) and (n_cos, n_ereco, n_cosz) as in the reference."""
    pv = view(p, t)
    osc = precomputed_osc if precomputed_osc is not None else compute_shared_osc_probs(p, t, fast, earth_uncertainty)
    bg_es, bg_cc = normalize_backgrounds(t, pv)
    es, cc = compute_oscillated_samples(t, pv, osc)
    w = np.asarray(t["exposure_weights"])
    wz = w[:, None]
    nz = len(w)
    incl, ang = _flag(t, "inclusive"), _flag(t, "angular")
    n_es = len(bg_es)
    if ang:
        ncos = int(t["n_cos"])
        r_ang, r_bg, amask = np.asarray(t["resp_es_angular"]), np.asarray(t["resp_bg_angular"]), np.asarray(t["angular_mask"])
    if es_mode:
        ne_d = apply_day_response(es["nue_day"], t["resp_es_nue"], t["eff_es_nue"])
        ne_n = apply_night_response(es["nue_night"], t["resp_es_nue"], t["eff_es_nue"])
        no_d = apply_day_response(es["nuother_day"], t["resp_es_nuother"], t["eff_es_nuother"])
        no_n = apply_night_response(es["nuother_night"], t["resp_es_nuother"], t["eff_es_nuother"])
        if ang:   # compute_ES_angular_event_rates, propagation_reco.jl:66-108
            sig_d, sig_n = ne_d + no_d, ne_n + no_n
            es_day = r_ang * sig_d[None, :] + r_bg * bg_es[None, :] * 0.5
            es_night = r_ang[:, :, None] * sig_n.T[None, :, :] + (r_bg[:, :, None] * (bg_es[:, None] * w[None, :])[None, :, :]) * 0.5
            es_day = es_day * amask[:, None]
            es_night = es_night * amask[:, None, None]
        else:     # compute_ES_event_rates, propagation_reco.jl:10-29
            es_day = (ne_d + no_d) + 0.5 * bg_es
            es_night = (ne_n + no_n) + 0.5 * (bg_es[None, :] * wz)
    else:
        es_day, es_night = (np.zeros((ncos, n_es)), np.zeros((ncos, n_es, nz))) if ang else (np.zeros(n_es), np.zeros((nz, n_es)))

    def cc_exclusive():   # compute_CC_event_rates, propagation_reco.jl:117-130
        if not cc_mode:
            return np.zeros(len(bg_cc)), np.zeros((nz, len(bg_cc)))
        return (apply_day_response(cc["day"], t["resp_cc"], t["eff_cc"]) + 0.5 * bg_cc,
                apply_night_response(cc["night"], t["resp_cc"], t["eff_cc"]) + 0.5 * (bg_cc[None, :] * wz))

    def cc_inclusive():   # compute_CC_inclusive_event_rates, propagation_reco.jl:139-170
        if not cc_mode:
            return (np.zeros((ncos, n_es)), np.zeros((ncos, n_es, nz))) if ang else (np.zeros(n_es), np.zeros((nz, n_es)))
        d = apply_day_response(cc["day"], t["resp_cc_incl"], t["eff_cc_incl"])
        n = apply_night_response(cc["night"], t["resp_cc_incl"], t["eff_cc_incl"])
        if ang:
            return r_bg * d[None, :], r_bg[:, :, None] * n.T[None, :, :]
        return d, n

    incl_spectrum = None
    if incl == 2:         # propagation_main.jl:109-137
        f_above, f_below = float(t["cc_f_above"]), float(t["cc_f_below"])
        ci_d, ci_n = cc_inclusive()
        incl_spectrum = f_above * ((ci_d.sum(axis=0) + ci_n.sum(axis=(0, 2))) if ang else (ci_d + ci_n.sum(axis=0)))
        es_day = es_day + f_above * ci_d
        es_night = es_night + f_above * ci_n
        cc_day, cc_night = cc_exclusive()
        cc_day, cc_night = cc_day * f_below, cc_night * f_below
        if es_mode:       # compute_ES_misID_CC_event_rates, propagation_reco.jl:187-199
            cc_day = cc_day + (apply_day_response(es["nue_day"], t["resp_cc_misid_nue"], t["eff_cc"]) +
                               apply_day_response(es["nuother_day"], t["resp_cc_misid_nuother"], t["eff_cc"]))
            cc_night = cc_night + (apply_night_response(es["nue_night"], t["resp_cc_misid_nue"], t["eff_cc"]) +
                                   apply_night_response(es["nuother_night"], t["resp_cc_misid_nuother"], t["eff_cc"]))
    elif incl == 1:       # propagation_main.jl:139-150
        ci_d, ci_n = cc_inclusive()
        incl_spectrum = (ci_d.sum(axis=0) + ci_n.sum(axis=(0, 2))) if ang else (ci_d + ci_n.sum(axis=0))
        es_day = es_day + ci_d
        es_night = es_night + ci_n
        cc_day, cc_night = np.zeros(len(bg_cc)), np.zeros((nz, len(bg_cc)))
    else:
        cc_day, cc_night = cc_exclusive()
    return dict(ES_day=es_day, CC_day=cc_day, ES_night=es_night, CC_night=cc_night,
                BG_ES_tot=bg_es, BG_CC_tot=bg_cc, CC_incl_spectrum=incl_spectrum)


# ---------------------------------------------------------------------------
# likelihood  (src/likelihoods/*.jl)
# ---------------------------------------------------------------------------
class NegativeRateError(RuntimeError):
    pass


def poisson_log_likelihood(n_expected, n_measured, flags: Optional[list] = None) -> float:
    """Soft-floored Poisson deviance, returned negated (likelihood_statistical.jl:1-28)."""
    n_expected = np.asarray(n_expected, dtype=np.float64)
    n_measured = np.asarray(n_measured, dtype=np.float64)
    if n_expected.shape != n_measured.shape:
        raise ValueError("Inputs must have the same length")
    llh = 0.0
    for raw_e, m in zip(n_expected.ravel(), n_measured.ravel()):
        if raw_e < 0:
            if raw_e > -1e-3:
                if flags is not None:
                    flags.append("warn")
            else:
                raise NegativeRateError(f"bin significantly negative (e={raw_e})")
        e = 0.5 * (raw_e + 1e-9 + math.sqrt((raw_e - 1e-9) ** 2 + 1e-18))
        llh += (e - m + m * (math.log(m) - math.log(e))) if m > 0 else e
    return -llh


def perbin_poisson_log_likelihood(n_expected, n_measured) -> np.ndarray:
    """likelihood_statistical.jl:31-57."""
    e = np.asarray(n_expected, dtype=np.float64)
    m = np.asarray(n_measured, dtype=np.float64)
    e_safe = 0.5 * (e + 1e-9 + np.sqrt((e - 1e-9) ** 2 + 1e-18))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(m > 0, -(e - m * np.log(e_safe)), -e)


def check_earth_norm_bounds(pv: ParamView) -> bool:
    """likelihood_debug.jl:1-4."""
    return bool(np.all((pv.earth_norms >= 0) & (pv.earth_norms <= 2)))


def llh_channel(rates_day, rates_night, obs_day, obs_night, idx: int, flags=None, night_only: bool = False) -> float:
    """Day vector plus every night row, bins idx:end (likelihood_builder.jl:5-20, 129-144).  night_only (the
    night-only Earth-density study of BASELINE config 4, solarosc_config.night_only) leaves the day vector out."""
    k = idx - 1
    out = 0.0 if night_only else poisson_log_likelihood(rates_day[k:], obs_day[k:], flags)
    for row, orow in zip(rates_night, obs_night):
        out += poisson_log_likelihood(row[k:], orow[k:], flags)
    return out


def llh_es_angle(rates_day, rates_night, obs_day, obs_night, idx: int, flags=None) -> float:
    """llh_ES_angle (likelihood_builder.jl:53-78): day (n_cos, n_E) and every night slice (n_cos, n_E, n_cosz),
    energy bins idx:end, all cos-scatter bins."""
    k = idx - 1
    out = poisson_log_likelihood(rates_day[:, k:].ravel(), obs_day[:, k:].ravel(), flags)
    for z in range(rates_night.shape[2]):
        out += poisson_log_likelihood(rates_night[:, k:, z].ravel(), obs_night[:, k:, z].ravel(), flags)
    return out


def total_llh(p, t, data, fast=True, earth_uncertainty=True, use_es=True, use_cc=True,
              precomputed_osc=None, es_mode=True, cc_mode=True, night_only=False) -> float:
    """likelihood_main.jl:8-35 + likelihood_builder.jl:216-259 for one detector: use_ES = ES_mode,
    use_CC = CC_mode && (!inclusive || semi_inclusive), llh_ES_angle when angular."""
    pv = view(p, t)
    if earth_uncertainty and not check_earth_norm_bounds(pv):
        return -math.inf
    r = propagate_samples(p, t, fast, earth_uncertainty, precomputed_osc, es_mode, cc_mode)
    llh = 0.0
    use_es = use_es and es_mode
    use_cc = use_cc and cc_mode and _flag(t, "inclusive") != 1
    if use_es and _flag(t, "angular"):
        llh += llh_es_angle(r["ES_day"], r["ES_night"], data["ES_day"], data["ES_night"], int(t["index_es"]))
    elif use_es:
        llh += llh_channel(r["ES_day"], r["ES_night"], data["ES_day"], data["ES_night"], int(t["index_es"]), night_only=night_only)
    if use_cc:
        llh += llh_channel(r["CC_day"], r["CC_night"], data["CC_day"], data["CC_night"], int(t["index_cc"]), night_only=night_only)
    return llh


def perbin_llh(p, t, data, fast=True, earth_uncertainty=True, rates=None):
    """make_perbin_likelihood with the Poisson per-bin terms (likelihood_builder.jl:22-51,146-174,261-333)."""
    pv = view(p, t)
    shapes = {k: np.shape(data[k]) for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
    if earth_uncertainty and not check_earth_norm_bounds(pv):
        return {k: np.full(s, -np.inf) for k, s in shapes.items()}
    r = rates if rates is not None else propagate_samples(p, t, fast, earth_uncertainty)
    out = {k: np.zeros(s) for k, s in shapes.items()}
    for ch, idx in (("ES", int(t["index_es"])), ("CC", int(t["index_cc"]))):
        n = shapes[ch + "_day"][0]
        sl = slice(idx - 1, n - 1)          # idx:n_bins-1, the overflow bin is excluded
        out[ch + "_day"][sl] = perbin_poisson_log_likelihood(r[ch + "_day"][sl], data[ch + "_day"][sl])
        out[ch + "_night"][:, sl] = perbin_poisson_log_likelihood(r[ch + "_night"][:, sl], data[ch + "_night"][:, sl])
    return out


def asimov_data(t, fast=True, earth_uncertainty=True):
    """Asimov data = rates at the true parameters (setup.jl:160-166, 212-217)."""
    r = propagate_samples(t["truth"], t, fast, earth_uncertainty)
    return {k: r[k] for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
