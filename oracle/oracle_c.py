"""ctypes binding of the C oracle (``oracle/solarosc_oracle.c``).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Mapping, Optional

import numpy as np

from dune_solaroscfitter_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsolarosc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "solarosc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
        tp, cp = C.POINTER(abi.Tables), C.POINTER(abi.Config)
        L.oracle_n_params.argtypes = [tp, cp]
        L.oracle_eval_llh.argtypes = [tp, cp, dp, C.c_size_t, dp, up, C.c_int]
        L.oracle_eval_oscprobs.argtypes = [tp, cp, dp, C.c_size_t, dp, dp]
        L.oracle_eval_rates.argtypes = [tp, cp, C.c_int, dp, C.c_size_t, dp, dp, dp, dp]
        L.oracle_poisson.argtypes = [dp, dp, C.c_size_t, up]
        L.oracle_poisson.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class COracle:
    """The C restatement bound to one set of tables (+ data)."""

    def __init__(self, tables: Mapping, data: Optional[Mapping] = None, fast: bool = True,
                 earth_uncertainty: bool = True, es_mode: bool = True, cc_mode: bool = True, night_only: bool = False):
        self.packed = abi.PackedTables(tables, data=None if data is None else [data], es_mode=es_mode, cc_mode=cc_mode)
        self.cfg = abi.make_config(fast=fast, earth_uncertainty=earth_uncertainty, night_only=night_only)
        self.P = lib().oracle_n_params(C.byref(self.packed.struct), C.byref(self.cfg))
        self.nrow = 1 + self.packed.n_cosz

    def _params(self, params):
        p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        assert p.shape[1] == self.P, (p.shape, self.P)
        return p

    def eval_llh(self, params, n_threads: int = 0):
        p = self._params(params)
        out = np.empty(len(p)); fl = np.zeros(len(p), dtype=np.uint32)
        lib().oracle_eval_llh(C.byref(self.packed.struct), C.byref(self.cfg), _p(p), len(p), _p(out),
                              fl.ctypes.data_as(C.POINTER(C.c_uint32)), n_threads)
        return out, fl

    def eval_oscprobs(self, params, want_p1e: bool = False):
        p = self._params(params)
        nT = self.packed.n_etrue
        osc = np.empty((len(p), 2, self.nrow, nT))
        p1e = np.empty((len(p), self.packed.n_paths, 2 * nT)) if want_p1e else None
        lib().oracle_eval_oscprobs(C.byref(self.packed.struct), C.byref(self.cfg), _p(p), len(p), _p(osc),
                                   _p(p1e) if want_p1e else None)
        return (osc, p1e) if want_p1e else osc

    def eval_rates(self, params, det: int = 0):
        p = self._params(params)
        nes, ncc = self.packed.n_ereco[det]
        es = np.empty((len(p), self.nrow, nes)); cc = np.empty((len(p), self.nrow, ncc))
        be = np.empty((len(p), nes)); bc = np.empty((len(p), ncc))
        lib().oracle_eval_rates(C.byref(self.packed.struct), C.byref(self.cfg), det, _p(p), len(p), _p(es), _p(cc),
                                _p(be), _p(bc))
        return dict(ES_day=es[:, 0], ES_night=es[:, 1:], CC_day=cc[:, 0], CC_night=cc[:, 1:], BG_ES_tot=be, BG_CC_tot=bc)

    @staticmethod
    def poisson(e, m):
        e = np.ascontiguousarray(e, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64)
        fl = C.c_uint32(0)
        return lib().oracle_poisson(_p(e), _p(m), e.size, C.byref(fl)), fl.value


def max_threads() -> int:
    L = lib()
    L.oracle_max_threads.restype = C.c_int
    return L.oracle_max_threads()
