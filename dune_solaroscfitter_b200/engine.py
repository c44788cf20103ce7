"""``SolarOscEngine``: one handle of the C-ABI library, batched over parameter points.

This is the Python face of ``include/solarosc_b200.h``; all arithmetic happens in the CUDA
kernels behind ``libsolarosc_b200.so``.  The per-call semantics mirror the reference functions
named in the header (``total_llh``, ``expected_rates``, ``make_perbin_likelihood``,
``compute_shared_osc_probs``, ``osc_prob_earth``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence

import numpy as np

from . import _lib, abi


class SolarOscError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"solarosc status {status}: {msg}")
        self.status = status


class NegativeRateError(SolarOscError):
    """Some expected rate was <= -1e-3 (the reference raises: likelihood_statistical.jl:9-14)."""


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class SolarOscEngine:
    def __init__(self, shared: Mapping, detectors: Optional[Sequence[Mapping]] = None,
                 data: Optional[Sequence[Optional[Mapping]]] = None, *, device: int = 0, fast: bool = True,
                 earth_uncertainty: bool = True, col_major: bool = False, max_batch: int = 0,
                 es_mode: bool = True, cc_mode: bool = True, det_names: Optional[Sequence[str]] = None,
                 night_only: bool = False):
        self._L = _lib.load()
        if data is not None and isinstance(data, Mapping):
            data = [data]
        self.packed = abi.PackedTables(shared, detectors, data, es_mode=es_mode, cc_mode=cc_mode, col_major=col_major)
        self.cfg = abi.make_config(device, fast, earth_uncertainty, col_major, max_batch, night_only)
        st = C.c_int(0)
        self._h = self._L.solarosc_create(C.byref(self.packed.struct), C.byref(self.cfg), C.byref(st))
        if not self._h:
            raise SolarOscError(st.value, (self._L.solarosc_last_error(None) or b"").decode())
        self.n_params = self._L.solarosc_n_params(self._h)
        self.param_names = self.packed.param_names(earth_uncertainty, det_names)
        self.col_major = bool(col_major)
        self.n_etrue, self.n_cosz, self.n_paths = self.packed.n_etrue, self.packed.n_cosz, self.packed.n_paths
        self.device = device
        self.rank, self.world = 0, 1

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.solarosc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers ----------------------------------------------------------------------------------
    def _check(self, status: int):
        if status == abi.OK:
            return
        msg = (self._L.solarosc_last_error(self._h) or b"").decode()
        if status == abi.ERR_NEGATIVE_RATE:
            raise NegativeRateError(status, msg)
        raise SolarOscError(status, msg)

    def _params(self, params) -> np.ndarray:
        p = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=np.float64)))
        if p.shape[1] != self.n_params:
            raise ValueError(f"parameter matrix must have {self.n_params} columns {self.param_names}, got {p.shape[1]}")
        return p

    def _night(self, n, n_ereco):
        return np.empty((n, n_ereco, self.n_cosz) if self.col_major else (n, self.n_cosz, n_ereco))

    def _es_arrays(self, n, det):
        """ES day / night outputs: with the angular reconstruction they carry the cos-scatter axis in the
        reference's index order (n_cos, n_ereco[, n_cosz]), reversed when col_major."""
        nes, nc = self.packed.n_ereco[det][0], self.packed.n_cos[det]
        if nc == 1:
            return np.empty((n, nes)), self._night(n, nes)
        if self.col_major:
            return np.empty((n, nes, nc)), np.empty((n, self.n_cosz, nes, nc))
        return np.empty((n, nc, nes)), np.empty((n, nc, nes, self.n_cosz))

    # -- API --------------------------------------------------------------------------------------
    def set_data(self, data: Mapping, det: int = 0):
        a = [np.ascontiguousarray(data[k], dtype=np.float64) for k in ("ES_day", "ES_night", "CC_day", "CC_night")]
        self._check(self._L.solarosc_set_data(self._h, det, *[_dp(x) for x in a]))

    def eval_llh(self, params, group: int = 1, return_flags: bool = False, raise_on_negative: bool = False,
                 out: Optional[np.ndarray] = None):
        p = self._params(params)
        if out is None:
            out = np.empty(len(p))
        elif out.shape != (len(p),) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array with one entry per point")
        fl = np.zeros(len(p), dtype=np.uint32) if return_flags else None      # the library takes NULL for "no flags"
        fp = fl.ctypes.data_as(C.POINTER(C.c_uint32)) if return_flags else None
        if group == 1:
            st = self._L.solarosc_eval_llh(self._h, _dp(p), len(p), _dp(out), fp)
        else:
            st = self._L.solarosc_eval_llh_grouped(self._h, _dp(p), len(p), group, _dp(out), fp)
        if st != abi.ERR_NEGATIVE_RATE or raise_on_negative:
            self._check(st)
        return (out, fl) if return_flags else out

    def eval_llh_grad(self, params, return_flags: bool = False):
        """lnL [n] and d lnL / d params [n, P] (what BAT's HMC asks of the log-density, mcmc.jl:102-103)."""
        p = self._params(params)
        out, grad = np.empty(len(p)), np.empty((len(p), self.n_params))
        fl = np.zeros(len(p), dtype=np.uint32)
        st = self._L.solarosc_eval_llh_grad(self._h, _dp(p), len(p), _dp(out), _dp(grad), fl.ctypes.data_as(C.POINTER(C.c_uint32)))
        if st != abi.ERR_NEGATIVE_RATE:
            self._check(st)
        return (out, grad, fl) if return_flags else (out, grad)

    def eval_rates(self, params, det: int = 0) -> Dict[str, np.ndarray]:
        p = self._params(params)
        n = len(p)
        nes, ncc = self.packed.n_ereco[det]
        es_day, es_night = self._es_arrays(n, det)
        r = dict(ES_day=es_day, ES_night=es_night, CC_day=np.empty((n, ncc)),
                 CC_night=self._night(n, ncc), BG_ES_tot=np.empty((n, nes)), BG_CC_tot=np.empty((n, ncc)))
        self._check(self._L.solarosc_eval_rates(self._h, det, _dp(p), n, _dp(r["ES_day"]), _dp(r["ES_night"]),
                                                _dp(r["CC_day"]), _dp(r["CC_night"]), _dp(r["BG_ES_tot"]), _dp(r["BG_CC_tot"])))
        return r

    def eval_cc_incl_spectrum(self, params, det: int = 0) -> Optional[np.ndarray]:
        """CC_incl_spectrum [n, n_ereco_es] (7th element of propagateSamples, propagation_main.jl:108-146); None for a
        detector of the exclusive analysis, where the reference returns ``nothing``."""
        if not self.packed.inclusive[det]:
            return None
        p = self._params(params)
        out = np.empty((len(p), self.packed.n_ereco[det][0]))
        self._check(self._L.solarosc_eval_cc_incl_spectrum(self._h, det, _dp(p), len(p), _dp(out)))
        return out

    def eval_perbin(self, params, det: int = 0) -> Dict[str, np.ndarray]:
        p = self._params(params)
        n = len(p)
        nes, ncc = self.packed.n_ereco[det]
        es_day, es_night = self._es_arrays(n, det)
        r = dict(ES_day=es_day, ES_night=es_night, CC_day=np.empty((n, ncc)), CC_night=self._night(n, ncc))
        self._check(self._L.solarosc_eval_perbin(self._h, det, _dp(p), n, _dp(r["ES_day"]), _dp(r["ES_night"]),
                                                 _dp(r["CC_day"]), _dp(r["CC_night"])))
        return r

    def eval_asymmetry(self, params, det: int = 0):
        """Per-point day-night asymmetries (ES, CC), reduced on the device."""
        p = self._params(params)
        a_es, a_cc = np.empty(len(p)), np.empty(len(p))
        self._check(self._L.solarosc_eval_asymmetry(self._h, det, _dp(p), len(p), _dp(a_es), _dp(a_cc)))
        return a_es, a_cc

    def eval_derived(self, params, weights, lref: Optional[Mapping] = None, det: int = 0) -> Dict[str, np.ndarray]:
        """Derived-mode block sums of derive_variables_from_chain.jl:260-444 for the (thinned-in) chain samples
        `params` with chain weights `weights`, reduced on the device.  `lref` holds the reference per-bin terms
        (keys ES_day, ES_night, CC_day, CC_night, e.g. eval_perbin at the truth).  Returns per-sample day rates and
        per-bin-term totals, and the weighted sums, all row-major ([.., 1 + n_cosz, n_ereco], row 0 = day).  With the
        angular ES reconstruction ES_day is the day spectrum summed over the cos-scatter bins and the ES sums come as
        mom_ES_day [6, n_cos, n_ereco] and mom_ES_night [6, n_cos, n_ereco, n_cosz] (lref in eval_perbin's shapes)."""
        p = self._params(params)
        n = len(p)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (n,):
            raise ValueError("weights must have one entry per sample")
        nes, ncc = self.packed.n_ereco[det]
        nz = self.n_cosz
        nc = self.packed.n_cos[det]

        def ref(ch, nr):
            if lref is None:
                return None
            night = np.asarray(lref[ch + "_night"], dtype=np.float64)
            if self.col_major:
                night = night.T
            return np.ascontiguousarray(np.vstack([np.asarray(lref[ch + "_day"], dtype=np.float64).reshape(1, nr), night.reshape(nz, nr)]))

        def ref_ang():
            """angular ES reference terms: [n_cos][n_ereco] day part, then [n_cos][n_ereco][n_cosz] night part"""
            if lref is None:
                return None
            day, night = np.asarray(lref["ES_day"], dtype=np.float64), np.asarray(lref["ES_night"], dtype=np.float64)
            if self.col_major:
                day, night = day.T, night.transpose(2, 1, 0)
            return np.ascontiguousarray(np.concatenate([day.reshape(nc * nes), night.reshape(nc * nes * nz)]))

        r_es, r_cc = (ref_ang() if nc > 1 else ref("ES", nes)), ref("CC", ncc)
        mom_es = np.empty((6, nc * nes * (1 + nz))) if nc > 1 else np.empty((6, 1 + nz, nes))
        out = dict(ES_day=np.empty((n, nes)), CC_day=np.empty((n, ncc)), ltot=np.empty((n, 4)),
                   mom_ES=mom_es, mom_CC=np.empty((6, 1 + nz, ncc)),
                   bg_mom_ES=np.empty((2, nes)), bg_mom_CC=np.empty((2, ncc)), scalars=np.empty(21))
        self._check(self._L.solarosc_eval_derived(self._h, det, _dp(p), n, _dp(w), _dp(r_es), _dp(r_cc), _dp(out["ES_day"]),
                                                  _dp(out["CC_day"]), _dp(out["ltot"]), _dp(out["mom_ES"]), _dp(out["mom_CC"]),
                                                  _dp(out["bg_mom_ES"]), _dp(out["bg_mom_CC"]), _dp(out["scalars"])))
        if nc > 1:   # angular: the six ES blocks split into their day [n_cos, n_ereco] and night [n_cos, n_ereco, n_cosz] parts
            out["mom_ES_day"] = mom_es[:, :nc * nes].reshape(6, nc, nes)
            out["mom_ES_night"] = mom_es[:, nc * nes:].reshape(6, nc, nes, nz)
        return out

    def eval_oscprobs(self, params) -> Dict[str, np.ndarray]:
        p = self._params(params)
        n, nT = len(p), self.n_etrue
        r = dict(nue_8B_day=np.empty((n, nT)), nue_8B_night=self._night(n, nT),
                 nue_hep_day=np.empty((n, nT)), nue_hep_night=self._night(n, nT))
        self._check(self._L.solarosc_eval_oscprobs(self._h, _dp(p), n, _dp(r["nue_8B_day"]), _dp(r["nue_8B_night"]),
                                                   _dp(r["nue_hep_day"]), _dp(r["nue_hep_night"])))
        for k in ("8B_day", "8B_night", "hep_day", "hep_night"):
            r["nuother_" + k] = 1 - r["nue_" + k]   # unitarity, propagation_osc.jl:140-144
        return r

    def eval_p1e(self, params) -> np.ndarray:
        p = self._params(params)
        shape = (len(p), 2 * self.n_etrue, self.n_paths) if self.col_major else (len(p), self.n_paths, 2 * self.n_etrue)
        out = np.empty(shape)
        self._check(self._L.solarosc_eval_p1e(self._h, _dp(p), len(p), _dp(out)))
        return out

    def eval_llh_device(self, d_params_ptr: int, n: int, d_lnl_ptr: int, d_flags_ptr: int = 0, group: int = 1,
                        stream_ptr: int = 0):
        """Enqueue on device-resident buffers (raw pointers, e.g. ``tensor.data_ptr()``); no synchronisation."""
        self._check(self._L.solarosc_eval_llh_device(self._h, C.c_void_p(d_params_ptr), n, group, C.c_void_p(d_lnl_ptr),
                                                     C.c_void_p(d_flags_ptr or None), C.c_void_p(stream_ptr or None)))

    # -- one process per GPU: NCCL communicator inside the library ----------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes (an ncclUniqueId) from rank 0, to be handed to every rank by the host's own means."""
        buf = C.create_string_buffer(abi.COMM_ID_BYTES)
        st = _lib.load().solarosc_comm_unique_id(buf)
        if st != abi.OK:
            raise SolarOscError(st, "solarosc_comm_unique_id failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Collective: every rank calls this with the same id (solarosc_comm_init)."""
        self._check(self._L.solarosc_comm_init(self._h, C.c_char_p(unique_id), rank, world))
        self.rank, self.world = rank, world

    def comm_init_from_torch(self):
        """Bootstrap the library's communicator over an initialised torch.distributed group: rank 0 draws the id and
        broadcasts its bytes (plumbing only; the gather itself is the library's own NCCL call)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        self.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)

    def comm_destroy(self):
        """Collective: every rank releases the communicator together (solarosc_comm_destroy).  A handle closed with its
        communicator still attached aborts it instead, without waiting for the other ranks."""
        self._check(self._L.solarosc_comm_destroy(self._h))

    def shard_bounds(self, n_total: int, group: int = 1, rank: Optional[int] = None, world: Optional[int] = None):
        lo, hi = C.c_size_t(0), C.c_size_t(0)
        self._check(self._L.solarosc_shard_bounds(n_total, group, self.rank if rank is None else rank,
                                                  self.world if world is None else world, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def eval_llh_sharded(self, params_local, n_total: int, group: int = 1, return_flags: bool = False, out=None, flags_out=None):
        """This rank's rows in, lnL of all n_total points out (all-gathered over NCCL inside the library).  `out` /
        `flags_out`: caller-owned result arrays (float64 / uint32, n_total entries; pinned memory makes the copy back
        asynchronous), allocated here when absent."""
        p = self._params(params_local) if len(params_local) else np.empty((0, self.n_params))
        out = np.empty(n_total) if out is None else out
        fl = flags_out if flags_out is not None else (np.empty(n_total, dtype=np.uint32) if return_flags else None)   # None: not copied back
        if out.shape != (n_total,) or out.dtype != np.float64 or (fl is not None and (fl.shape != (n_total,) or fl.dtype != np.uint32)):
            raise ValueError("out / flags_out must be float64 / uint32 arrays of n_total entries")
        st = self._L.solarosc_eval_llh_sharded(self._h, _dp(p), n_total, group, _dp(out),
                                               fl.ctypes.data_as(C.POINTER(C.c_uint32)) if fl is not None else None)
        if st != abi.ERR_NEGATIVE_RATE:
            self._check(st)
        return (out, fl) if return_flags else out

    def eval_llh_sharded_device(self, d_params_local_ptr: int, n_total: int, d_lnl_all_ptr: int, d_flags_all_ptr: int = 0,
                                group: int = 1, stream_ptr: int = 0):
        self._check(self._L.solarosc_eval_llh_sharded_device(self._h, C.c_void_p(d_params_local_ptr or None), n_total, group,
                                                             C.c_void_p(d_lnl_all_ptr), C.c_void_p(d_flags_all_ptr or None),
                                                             C.c_void_p(stream_ptr or None)))

    def poisson_llh(self, expected, measured):
        e = np.ascontiguousarray(np.atleast_2d(expected), dtype=np.float64)
        m = np.ascontiguousarray(np.atleast_2d(measured), dtype=np.float64)
        if e.shape != m.shape:
            raise ValueError("Inputs must have the same length")   # likelihood_statistical.jl:2-4
        out = np.empty(len(e))
        fl = np.zeros(len(e), dtype=np.uint32)
        self._check(self._L.solarosc_poisson_llh(self._h, _dp(e), _dp(m), e.shape[0], e.shape[1], _dp(out),
                                                 fl.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def batcher(self, n_participants: int, want_grad: bool = False, max_wait_ms: int = 0) -> "Batcher":
        """Rendezvous for n single-point callers (one per chain thread): see include/solarosc_b200.h."""
        return Batcher(self, n_participants, want_grad, max_wait_ms)

    def set_timing(self, enabled: bool = True):
        self._check(self._L.solarosc_set_timing(self._h, int(enabled)))

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0)
        self._check(self._L.solarosc_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def last_timing(self):
        a, b = C.c_float(0), C.c_float(0)
        n = self._L.solarosc_last_timing(self._h, C.byref(a), C.byref(b))
        return dict(ms_osc=a.value, ms_rates=b.value, launches=n, fused=bool(self._L.solarosc_last_fused(self._h)))


def eval_llh_multi(engines: Sequence["SolarOscEngine"], params, group: int = 1, return_flags: bool = False):
    """One batch dealt over several engines (one per GPU, same tables) from this process: contiguous slices of whole
    groups, evaluated concurrently, lnL of every slice written straight into one result array
    (solarosc_eval_llh_multi)."""
    e0 = engines[0]
    p = e0._params(params)
    out = np.empty(len(p))
    fl = np.zeros(len(p), dtype=np.uint32)
    hs = (C.c_void_p * len(engines))(*[e._h for e in engines])
    st = e0._L.solarosc_eval_llh_multi(hs, len(engines), _dp(p), len(p), group, _dp(out), fl.ctypes.data_as(C.POINTER(C.c_uint32)))
    if st != abi.ERR_NEGATIVE_RATE:
        e0._check(st)
    return (out, fl) if return_flags else out


class Batcher:
    """N threads each call ``eval(slot, params)`` with ONE point; one batched launch serves them all."""

    def __init__(self, engine: SolarOscEngine, n_participants: int, want_grad: bool = False, max_wait_ms: int = 0):
        self._L, self._engine = engine._L, engine   # the handle must outlive the batcher
        self.n_params, self.want_grad = engine.n_params, bool(want_grad)
        self._b = self._L.solarosc_batcher_create(engine._h, n_participants, int(want_grad), max_wait_ms)
        if not self._b:
            raise SolarOscError(abi.ERR_ARG, "solarosc_batcher_create failed")

    def eval(self, slot: int, params):
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(self.n_params)
        lnl, fl = C.c_double(), C.c_uint32()
        g = np.empty(self.n_params) if self.want_grad else None
        st = self._L.solarosc_batcher_eval(self._b, slot, _dp(p), C.byref(lnl), _dp(g), C.byref(fl))
        if st not in (abi.OK, abi.ERR_NEGATIVE_RATE):
            raise SolarOscError(st, "solarosc_batcher_eval failed")
        return (lnl.value, g) if self.want_grad else lnl.value

    def leave(self, slot: int):
        self._L.solarosc_batcher_leave(self._b, slot)

    def stats(self):
        nb, npnt = C.c_ulonglong(), C.c_ulonglong()
        self._L.solarosc_batcher_stats(self._b, C.byref(nb), C.byref(npnt))
        return nb.value, npnt.value

    def close(self):
        if getattr(self, "_b", None):
            self._L.solarosc_batcher_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
