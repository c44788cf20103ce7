"""Host-side mirror of the reference's call surface for the hot path, backed by the CUDA engine.

The reference is a script-style Julia application: ``setup.jl`` fills module-level globals and the
per-evaluation functions read them.  A Julia host keeps those scripts and swaps the arithmetic for
``ccall``s into ``libsolarosc_b200.so`` (see ``julia/b200_interface.jl`` and INTEGRATION.md).  No
Julia exists in this image, so this module restates the same surface in Python -- same function
names, argument meaning, return shapes and error behaviour -- with a ``SolarOscSession`` standing
in for the globals:

    reference (Julia)                                          here
    ---------------------------------------------------------  -----------------------------------
    compute_shared_osc_probs(params, ssm)      propagation_main.jl:47-54    session.compute_shared_osc_probs
    propagateSamples(unosc, resp, params, ...) propagation_main.jl:56-157   session.propagateSamples
    expected_rates(d, params; precomputed_osc) likelihood_core.jl:17-35     expected_rates
    make_likelihood(d; ...)(params; ...)       likelihood_builder.jl:216-259 make_likelihood
    make_perbin_likelihood(d; ...)(params)     likelihood_builder.jl:261-333 make_perbin_likelihood
    total_llh(params)                          likelihood_main.jl:30-35     session.total_llh
    poissonLogLikelihood(e, m)                 likelihood_statistical.jl:1-28 session.poissonLogLikelihood
    check_earth_norm_bounds(params)            likelihood_debug.jl:1-4      check_earth_norm_bounds

plus the batched siblings the run-mode drivers need (``total_llh_batch``, ``expected_rates_batch``).
Parameters are mappings keyed by the reference's names; families (``earth_norm_*``,
``{det}_ES_bg_norm_*``, ``{det}_CC_bg_norm_*``) are gathered by prefix and ordered by the names
sorted AS STRINGS, exactly like ``_params_by_prefix`` (core.jl:6-15).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Callable, Dict, Mapping, Optional, Sequence

import numpy as np

from . import abi
from .engine import NegativeRateError, SolarOscEngine


def params_by_prefix(params: Mapping[str, Any], prefix: str):
    """Values of every field whose name starts with ``prefix``, names sorted as strings (core.jl:6-15)."""
    names = sorted(k for k in params.keys() if str(k).startswith(prefix))
    return tuple(params[k] for k in names)


def check_earth_norm_bounds(params: Mapping[str, Any]) -> bool:
    """likelihood_debug.jl:1-4."""
    return all(0 <= x <= 2 for x in params_by_prefix(params, "earth_norm_"))


@dataclass
class LikelihoodInputs:
    """likelihood_core.jl:2-15 (``f`` is the propagation closure of setup.jl:153-157)."""
    nObserved: Mapping[str, np.ndarray]
    energies: Any
    Mreco: Any
    SSM: Any
    MC_no_osc: Any
    BG: Any
    f: Callable
    ES_mode: bool
    CC_mode: bool
    index_ES: int
    index_CC: int
    det_name: str


class SolarOscSession:
    """The reference's set-up globals plus one engine handle.

    ``shared``/``detectors`` use the field names of ``include/solarosc_b200.h``; ``det_names`` are the
    detector names that prefix the background-norm parameters (``HD_CC_bg_norm_1`` ...).
    """

    def __init__(self, shared: Mapping, detectors: Optional[Sequence[Mapping]] = None,
                 det_names: Sequence[str] = ("HD",), data: Optional[Sequence[Optional[Mapping]]] = None, *,
                 fast: bool = True, earth_uncertainty: bool = True, cc_xsec_norm_true: float = 1.0,
                 device: int = 0, max_batch: int = 0, es_mode: bool = True, cc_mode: bool = True):
        self.det_names = list(det_names)
        self.earth_uncertainty = earth_uncertainty
        self.cc_xsec_norm_true = float(cc_xsec_norm_true)
        self.engine = SolarOscEngine(shared, detectors, data, device=device, fast=fast,
                                     earth_uncertainty=earth_uncertainty, max_batch=max_batch,
                                     es_mode=es_mode, cc_mode=cc_mode, det_names=self.det_names)
        self.n_layers = self.engine.packed.n_layers
        self.n_norms = self.engine.packed.n_norms
        self._has_data = data is not None

    def close(self):
        self.engine.close()

    # -- parameter packing ----------------------------------------------------------------------
    def pack(self, params: Mapping[str, Any]) -> np.ndarray:
        """One row of the C-ABI parameter matrix from a name -> value mapping.

        Missing ``cc_xsec_*`` fall back to (cc_xsec_norm_true, 0, 0) (propagation_osc.jl:185-187);
        fields the path does not read (``ES_asymmetry`` ...) are ignored (setup.jl:117-118).
        """
        row = [params["sin2_th12"], params["sin2_th13"], params["dm2_21"],
               params["integrated_8B_flux"], params["integrated_HEP_flux"],
               params.get("cc_xsec_norm", self.cc_xsec_norm_true), params.get("cc_xsec_tilt", 0.0),
               params.get("cc_xsec_curv", 0.0)]
        if self.earth_uncertainty:
            norms = params_by_prefix(params, "earth_norm_")
            if len(norms) != self.n_layers:
                raise KeyError(f"expected {self.n_layers} earth_norm_* parameters, got {len(norms)}")
            row += list(norms)
        for name, (n_es, n_cc) in zip(self.det_names, self.n_norms):
            es = params_by_prefix(params, f"{name}_ES_bg_norm_")
            cc = params_by_prefix(params, f"{name}_CC_bg_norm_")
            if len(es) < min(n_es, 1) or len(cc) < n_cc:
                raise KeyError(f"missing background norms for detector {name}")
            # only ES_bg_norm_1 is read (propagation_bg.jl:56-57); the others keep their slot
            row += list(es[:n_es]) + [es[0] if es else 0.0] * max(0, n_es - len(es))
            row += list(cc[:n_cc])
        return np.asarray(row, dtype=np.float64)

    def pack_many(self, params) -> np.ndarray:
        if isinstance(params, np.ndarray):
            return np.atleast_2d(params)
        if isinstance(params, Mapping):
            return self.pack(params)[None, :]
        return np.stack([self.pack(p) for p in params])

    # -- oscillation stage --------------------------------------------------------------------------
    def compute_shared_osc_probs(self, params, solarModel=None) -> Dict[str, np.ndarray]:
        """8-field result of propagation_osc.jl:146-155 (night arrays are (n_cosz, n_Etrue))."""
        r = self.engine.eval_oscprobs(self.pack_many(params))
        return {k: v[0] for k, v in r.items()}

    def osc_prob_earth(self, params) -> np.ndarray:
        """p1e[ip, ie] of Osc.NumOsc.Fast.osc_prob_earth (osc.jl:448-477)."""
        return self.engine.eval_p1e(self.pack_many(params))[0]

    # -- event rates --------------------------------------------------------------------------------------
    def propagateSamples(self, unoscillatedSample=None, responseMatrices=None, params=None, solarModel=None,
                         bin_edges=None, raw_backgrounds=None, det_flags=None, *, precomputed_osc=None,
                         verbose=False):
        """7-tuple of propagation_main.jl:155-156 (every branch; CC_incl_spectrum is None in the exclusive analysis).

        The table arguments are accepted for signature compatibility; the tables in use are the ones
        the session was created with.  ``precomputed_osc`` is accepted and not needed: the oscillation
        stage is a pure function of ``params`` and runs on the device in the same call.
        """
        det = 0
        if det_flags is not None:
            name = det_flags["det_name"] if isinstance(det_flags, Mapping) else getattr(det_flags, "det_name")
            det = self.det_names.index(name)
        pm = self.pack_many(params)
        r = self.engine.eval_rates(pm, det)
        spec = self.engine.eval_cc_incl_spectrum(pm, det)
        return (r["ES_day"][0], r["CC_day"][0], r["ES_night"][0], r["CC_night"][0],
                r["BG_ES_tot"][0], r["BG_CC_tot"][0], None if spec is None else spec[0])

    def expected_rates_batch(self, params, det: int = 0) -> Dict[str, np.ndarray]:
        return self.engine.eval_rates(self.pack_many(params), det)

    # -- likelihood ------------------------------------------------------------------------------------------
    def set_data(self, data: Mapping, det: int = 0):
        self.engine.set_data(data, det)
        self._has_data = True

    def asimov(self, true_params: Mapping[str, Any]):
        """Asimov data = propagateSamples at the true parameters, per detector (setup.jl:160-166)."""
        out = []
        for d in range(len(self.det_names)):
            r = self.engine.eval_rates(self.pack_many(true_params), d)
            data = {k: r[k][0] for k in ("ES_day", "ES_night", "CC_day", "CC_night")}
            self.engine.set_data(data, d)
            out.append(data)
        self._has_data = True
        return out

    def total_llh(self, params) -> float:
        """likelihood_main.jl:30-35.  Raises like the reference on a significantly negative rate."""
        return float(self.total_llh_batch(params)[0])

    def total_llh_batch(self, params, group: int = 1) -> np.ndarray:
        return self.engine.eval_llh(self.pack_many(params), group=group, raise_on_negative=True)

    def total_llh_and_grad_batch(self, params):
        """lnL [n] and d lnL / d params [n, P]: the value + gradient BAT's HMC takes from the log-density where the
        reference differentiates total_llh with ForwardDiff (mcmc.jl:102-103)."""
        return self.engine.eval_llh_grad(self.pack_many(params))

    def total_llh_and_grad(self, params):
        lnl, grad = self.total_llh_and_grad_batch([params])
        return float(lnl[0]), dict(zip(self.engine.param_names, grad[0]))

    def per_bin_llh(self, params, det: int = 0) -> Dict[str, np.ndarray]:
        r = self.engine.eval_perbin(self.pack_many(params), det)
        return {k: v[0] for k, v in r.items()}

    def poissonLogLikelihood(self, nExpected, nMeasured) -> float:
        e, m = np.asarray(nExpected, dtype=np.float64), np.asarray(nMeasured, dtype=np.float64)
        if e.shape != m.shape:
            raise ValueError("Inputs must have the same length")
        return float(self.engine.poisson_llh(e.ravel(), m.ravel())[0])

    def likelihood_inputs(self, det: int = 0, nObserved: Optional[Mapping] = None) -> LikelihoodInputs:
        """The struct setup.jl:272-285 builds, with ``f`` bound to this session."""
        name = self.det_names[det]
        dt = self.engine.packed.det_array[det]
        flags = {"det_name": name}

        def prop_fn(unosc, resp, pars, ssm, edges, bg, precomputed_osc=None):
            return self.propagateSamples(unosc, resp, pars, ssm, edges, bg, flags, precomputed_osc=precomputed_osc)

        return LikelihoodInputs(nObserved or {}, None, None, None, None, None, prop_fn, bool(dt.es_mode),
                                bool(dt.cc_mode), int(dt.index_es), int(dt.index_cc), name)


def expected_rates(d: LikelihoodInputs, parameters, precomputed_osc=None) -> Dict[str, np.ndarray]:
    """likelihood_core.jl:17-35."""
    es_d, cc_d, es_n, cc_n, bg_es, bg_cc, _ = d.f(d.MC_no_osc, d.Mreco, parameters, d.SSM, d.energies, d.BG,
                                                  precomputed_osc=precomputed_osc)
    return dict(ES_day=es_d, CC_day=cc_d, ES_night=es_n, CC_night=cc_n, BG_ES_tot=bg_es, BG_CC_tot=bg_cc)


def make_likelihood(session: SolarOscSession, d: LikelihoodInputs, use_ES: bool = True, use_CC: bool = True):
    """likelihood_builder.jl:216-259 for the Poisson ES/CC terms, evaluated on the device."""
    det = session.det_names.index(d.det_name)
    dt = session.engine.packed.det_array[det]
    if bool(dt.es_mode) != bool(use_ES and d.ES_mode) or bool(dt.cc_mode) != bool(use_CC and d.CC_mode):
        raise ValueError("channel selection is fixed at session creation (es_mode / cc_mode)")

    def llh(parameters, precomputed_osc=None):
        if not check_earth_norm_bounds(parameters):
            return -math.inf
        return session.total_llh(parameters)

    return llh


def make_perbin_likelihood(session: SolarOscSession, d: LikelihoodInputs):
    """likelihood_builder.jl:261-333."""
    det = session.det_names.index(d.det_name)

    def perbin(parameters, rates=None):
        return session.per_bin_llh(parameters, det)

    return perbin


__all__ = ["SolarOscSession", "LikelihoodInputs", "expected_rates", "make_likelihood", "make_perbin_likelihood",
           "check_earth_norm_bounds", "params_by_prefix", "NegativeRateError", "abi"]
