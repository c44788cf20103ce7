"""ctypes mirror of ``include/solarosc_b200.h`` (struct layouts and constants only).

Keeps numpy arrays alive for as long as the packed structs that point into them.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np

ABI_VERSION = 3
COMM_ID_BYTES = 128
MAX_LAYERS = 8
MAX_BG = 8
N_FIXED_PARAMS = 8

OK, ERR_ARG, ERR_CUDA, ERR_ALLOC, ERR_NEGATIVE_RATE = 0, 1, 2, 3, 4
FLAG_EARTH_OOB, FLAG_NEG_WARN, FLAG_NEG_ERROR, FLAG_NONFINITE = 1, 2, 4, 8

PARAM_NAMES_FIXED = ("sin2_th12", "sin2_th13", "dm2_21", "integrated_8B_flux", "integrated_HEP_flux",
                     "cc_xsec_norm", "cc_xsec_tilt", "cc_xsec_curv")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class DetectorExt(C.Structure):
    """solarosc_detector_ext: tables of the inclusive / semi-inclusive / angular branches."""
    _fields_ = [
        ("inclusive", C.c_int32), ("angular", C.c_int32), ("n_cos", C.c_int32), ("reserved", C.c_int32),
        ("cc_f_above", C.c_double), ("cc_f_below", C.c_double),
        ("resp_cc_incl", _dp), ("eff_cc_incl", _dp), ("resp_cc_misid_nue", _dp), ("resp_cc_misid_nuother", _dp),
        ("resp_es_angular", _dp), ("resp_bg_angular", _dp), ("angular_mask", _dp),
    ]


class Detector(C.Structure):
    _fields_ = [
        ("es_mode", C.c_int32), ("cc_mode", C.c_int32),
        ("n_ereco_es", C.c_int32), ("n_ereco_cc", C.c_int32),
        ("index_es", C.c_int32), ("index_cc", C.c_int32),
        ("unosc_es_nue_8b", _dp), ("unosc_es_nuother_8b", _dp), ("unosc_cc_8b", _dp),
        ("unosc_es_nue_hep", _dp), ("unosc_es_nuother_hep", _dp), ("unosc_cc_hep", _dp),
        ("log_e_norm", _dp),
        ("resp_es_nue", _dp), ("resp_es_nuother", _dp), ("resp_cc", _dp),
        ("eff_es_nue", _dp), ("eff_es_nuother", _dp), ("eff_cc", _dp),
        ("n_bg_es", C.c_int32), ("n_bg_cc", C.c_int32),
        ("raw_bg_es", _dp), ("raw_bg_cc", _dp),
        ("bg_es_sides", _ip), ("bg_es_par", _ip), ("bg_cc_par", _ip),
        ("obs_es_day", _dp), ("obs_es_night", _dp), ("obs_cc_day", _dp), ("obs_cc_night", _dp),
        ("ext", C.POINTER(DetectorExt)),
    ]


class Tables(C.Structure):
    _fields_ = [
        ("n_etrue", C.c_int32), ("n_cosz", C.c_int32), ("n_cosz_sub", C.c_int32),
        ("n_layers", C.c_int32), ("n_solar", C.c_int32), ("n_det", C.c_int32),
        ("e_calc", _dp), ("seg_start", _ip), ("seg_layer", _ip), ("seg_length", _dp),
        ("earth_lookup", _dp), ("exposure_weights", _dp),
        ("avg_ne_8b", C.c_double), ("avg_ne_hep", C.c_double),
        ("solar_ne", _dp), ("solar_prod_8b", _dp), ("solar_prod_hep", _dp),
        ("cc_xsec_norm_true", C.c_double),
        ("det", C.POINTER(Detector)),
    ]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("fast", C.c_int32),
                ("earth_uncertainty", C.c_int32), ("col_major", C.c_int32), ("max_batch", C.c_int32),
                ("night_only", C.c_int32), ("reserved", C.c_int32)]


_DET_F64 = ("unosc_es_nue_8b", "unosc_es_nuother_8b", "unosc_cc_8b", "unosc_es_nue_hep",
            "unosc_es_nuother_hep", "unosc_cc_hep", "log_e_norm", "resp_es_nue", "resp_es_nuother",
            "resp_cc", "eff_es_nue", "eff_es_nuother", "eff_cc", "raw_bg_es", "raw_bg_cc")
_DET_I32 = ("bg_es_sides", "bg_es_par", "bg_cc_par")
_EXT_F64 = ("resp_cc_incl", "eff_cc_incl", "resp_cc_misid_nue", "resp_cc_misid_nuother", "resp_es_angular",
            "resp_bg_angular", "angular_mask")
_OBS = (("obs_es_day", "ES_day"), ("obs_es_night", "ES_night"), ("obs_cc_day", "CC_day"), ("obs_cc_night", "CC_night"))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a: Optional[np.ndarray], typ):
    return typ() if a is None else a.ctypes.data_as(typ)


class PackedTables:
    """A ``solarosc_tables`` struct plus the arrays it points to.

    ``shared`` holds the shared tables; ``detectors`` is a list of per-detector mappings using the
    field names of the header (a single mapping holding both kinds of keys, like the golden
    ``.npz``, is accepted as one detector).  ``data`` optionally supplies observed rates per
    detector (``ES_day`` / ``ES_night`` / ``CC_day`` / ``CC_night``).
    """

    def __init__(self, shared: Mapping[str, Any], detectors: Optional[Sequence[Mapping[str, Any]]] = None,
                 data: Optional[Sequence[Optional[Mapping[str, Any]]]] = None,
                 es_mode: bool = True, cc_mode: bool = True, col_major: bool = False):
        # col_major: 2-D arrays are numpy views of Julia memory, i.e. their shapes are reversed
        # ((n_ereco, n_etrue) for a response matrix, (n_ereco, n_cosz) for night data)
        if detectors is None:
            detectors = [shared]
        self._keep: List[np.ndarray] = []
        k = self._keep
        n_det = len(detectors)
        self.det_array = (Detector * n_det)()
        self.ext_array = (DetectorExt * n_det)()
        self.n_norms = []
        self.n_cos = []
        self.inclusive = []
        for d, src in enumerate(detectors):
            ds = self.det_array[d]
            ds.es_mode = int(src.get("es_mode", es_mode)) if hasattr(src, "get") else int(es_mode)
            ds.cc_mode = int(src.get("cc_mode", cc_mode)) if hasattr(src, "get") else int(cc_mode)
            for name in _DET_F64:
                a = _f64(src[name]); k.append(a); setattr(ds, name, _ptr(a, _dp))
            for name in _DET_I32:
                a = _i32(src[name]); k.append(a); setattr(ds, name, _ptr(a, _ip))
            ax = 0 if col_major else 1
            ds.n_ereco_es = int(np.shape(src["resp_es_nue"])[ax])
            ds.n_ereco_cc = int(np.shape(src["resp_cc"])[ax])
            ds.index_es, ds.index_cc = int(src["index_es"]), int(src["index_cc"])
            ds.n_bg_es, ds.n_bg_cc = int(np.shape(src["raw_bg_es"])[0]), int(np.shape(src["raw_bg_cc"])[0])
            obs = None if data is None else data[d]
            for cname, key in _OBS:
                if obs is None:
                    setattr(ds, cname, _dp())
                else:
                    a = _f64(obs[key]); k.append(a); setattr(ds, cname, _ptr(a, _dp))
            self.n_norms.append((int(np.count_nonzero(src["bg_es_par"])), int(np.count_nonzero(src["bg_cc_par"]))))
            # optional keys `inclusive` (0 / 1 / 2 semi-inclusive) and `angular` select the other analysis branches
            incl = int(src["inclusive"]) if "inclusive" in src else 0
            ang = int(src["angular"]) if "angular" in src else 0
            self.n_cos.append(int(src["n_cos"]) if ang else 1)
            self.inclusive.append(incl)
            if incl or ang:
                x = self.ext_array[d]
                x.inclusive, x.angular, x.n_cos = incl, ang, int(src["n_cos"]) if ang else 0
                x.cc_f_above = float(src["cc_f_above"]) if incl == 2 else 1.0
                x.cc_f_below = float(src["cc_f_below"]) if incl == 2 else 1.0
                for name in _EXT_F64:
                    if name in src and src[name] is not None:
                        a = _f64(src[name]); k.append(a); setattr(x, name, _ptr(a, _dp))
                ds.ext = C.pointer(x)
        t = self.struct = Tables()
        t.n_etrue = int(len(shared["e_calc"]) // 2)
        t.n_cosz = int(len(shared["exposure_weights"]))
        t.n_cosz_sub = int(shared["n_cosz_sub"])
        t.n_layers = int(len(shared["earth_lookup"]))
        t.n_det = n_det
        for name in ("e_calc", "seg_length", "earth_lookup", "exposure_weights"):
            a = _f64(shared[name]); k.append(a); setattr(t, name, _ptr(a, _dp))
        for name in ("seg_start", "seg_layer"):
            a = _i32(shared[name]); k.append(a); setattr(t, name, _ptr(a, _ip))
        t.avg_ne_8b, t.avg_ne_hep = float(shared["avg_ne_8b"]), float(shared["avg_ne_hep"])
        if "solar_ne" in shared and shared["solar_ne"] is not None:
            t.n_solar = int(len(shared["solar_ne"]))
            for name in ("solar_ne", "solar_prod_8b", "solar_prod_hep"):
                a = _f64(shared[name]); k.append(a); setattr(t, name, _ptr(a, _dp))
        else:
            t.n_solar = 0
        t.cc_xsec_norm_true = float(shared.get("cc_xsec_norm_true", 1.0)) if hasattr(shared, "get") else 1.0
        t.det = C.cast(self.det_array, C.POINTER(Detector))
        self.n_layers, self.n_det = t.n_layers, n_det
        self.n_etrue, self.n_cosz, self.n_paths = t.n_etrue, t.n_cosz, t.n_cosz * t.n_cosz_sub
        self.n_ereco = [(self.det_array[d].n_ereco_es, self.det_array[d].n_ereco_cc) for d in range(n_det)]

    def n_params(self, earth_uncertainty: bool = True) -> int:
        return N_FIXED_PARAMS + (self.n_layers if earth_uncertainty else 0) + sum(a + b for a, b in self.n_norms)

    def param_names(self, earth_uncertainty: bool = True, det_names: Optional[Sequence[str]] = None) -> List[str]:
        names = list(PARAM_NAMES_FIXED)
        if earth_uncertainty:
            names += [f"earth_norm_{i + 1}" for i in range(self.n_layers)]
        for d, (ne, nc) in enumerate(self.n_norms):
            dn = det_names[d] if det_names else f"det{d}"
            names += [f"{dn}_ES_bg_norm_{i + 1}" for i in range(ne)]
            names += [f"{dn}_CC_bg_norm_{i + 1}" for i in range(nc)]
        return names


def make_config(device: int = 0, fast: bool = True, earth_uncertainty: bool = True, col_major: bool = False,
                max_batch: int = 0, night_only: bool = False) -> Config:
    return Config(ABI_VERSION, int(device), int(fast), int(earth_uncertainty), int(col_major), int(max_batch),
                  int(night_only), 0)
