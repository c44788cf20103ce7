"""Loader of ``libsolarosc_b200.so`` (the C-ABI library holding the sm_100a kernels).

There is no CPU or PyTorch fallback anywhere in this package: if the shared library is missing
this raises, and if no B200 is present ``solarosc_create`` fails with ``SOLAROSC_ERR_CUDA``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import abi

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("SOLAROSC_LIB") or os.path.join(_PKG, "libsolarosc_b200.so")
_SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("solarosc_b200.cu", "solarosc_device.cuh", "solarosc_kernels.cuh", "solarosc_pipe.cuh", "solarosc_grad.cuh", "solarosc_adjoint.cuh", "solarosc_adjoint_math.cuh", "solarosc_general.cuh", "solarosc_setup.cuh", "solarosc_batcher.hpp", "solarosc_comm.hpp")]
_SOURCES.append(os.path.join(_ROOT, "include", "solarosc_b200.h"))

EXPORTS = (
    "solarosc_abi_version", "solarosc_create", "solarosc_destroy", "solarosc_last_error", "solarosc_n_params",
    "solarosc_set_data", "solarosc_eval_llh", "solarosc_eval_llh_grouped", "solarosc_eval_llh_multi", "solarosc_eval_llh_grad", "solarosc_eval_rates", "solarosc_eval_cc_incl_spectrum",
    "solarosc_eval_perbin", "solarosc_eval_asymmetry", "solarosc_eval_derived", "solarosc_eval_oscprobs", "solarosc_eval_p1e", "solarosc_eval_llh_device",
    "solarosc_poisson_llh", "solarosc_set_timing", "solarosc_last_timing", "solarosc_last_fused", "solarosc_measure_fp64_peak",
    "solarosc_comm_unique_id", "solarosc_comm_init", "solarosc_comm_destroy", "solarosc_shard_bounds", "solarosc_eval_llh_sharded",
    "solarosc_eval_llh_sharded_device",
    "solarosc_batcher_create", "solarosc_batcher_create_with", "solarosc_batcher_eval", "solarosc_batcher_leave",
    "solarosc_batcher_stats", "solarosc_batcher_destroy",
    "solarosc_setup_bin_index", "solarosc_setup_bin_index_device", "solarosc_setup_histogram",
    "solarosc_setup_response_matrix", "solarosc_setup_earth_paths",
)

BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_uint32))

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    stale = not os.path.exists(LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in _SOURCES)
    if force or stale:
        cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC,-O3", "-shared", "-o", LIB_PATH, _SOURCES[0], "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libsolarosc_b200.so")
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                           "this package has no fallback implementation")
    L = C.CDLL(LIB_PATH)
    dp, up, sz = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_size_t
    H = C.c_void_p
    L.solarosc_abi_version.restype = C.c_int
    L.solarosc_create.argtypes = [C.POINTER(abi.Tables), C.POINTER(abi.Config), C.POINTER(C.c_int)]
    L.solarosc_create.restype = H
    L.solarosc_destroy.argtypes = [H]
    L.solarosc_destroy.restype = None
    L.solarosc_last_error.argtypes = [H]
    L.solarosc_last_error.restype = C.c_char_p
    L.solarosc_n_params.argtypes = [H]
    L.solarosc_set_data.argtypes = [H, C.c_int, dp, dp, dp, dp]
    L.solarosc_eval_llh.argtypes = [H, dp, sz, dp, up]
    L.solarosc_eval_llh_grouped.argtypes = [H, dp, sz, sz, dp, up]
    L.solarosc_eval_llh_grad.argtypes = [H, dp, sz, dp, dp, up]
    L.solarosc_eval_rates.argtypes = [H, C.c_int, dp, sz, dp, dp, dp, dp, dp, dp]
    L.solarosc_eval_perbin.argtypes = [H, C.c_int, dp, sz, dp, dp, dp, dp]
    L.solarosc_eval_cc_incl_spectrum.argtypes = [H, C.c_int, dp, sz, dp]
    L.solarosc_comm_unique_id.argtypes = [C.c_void_p]
    L.solarosc_comm_init.argtypes = [H, C.c_void_p, C.c_int, C.c_int]
    L.solarosc_comm_destroy.argtypes = [H]
    L.solarosc_shard_bounds.argtypes = [sz, sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    L.solarosc_eval_llh_sharded.argtypes = [H, dp, sz, sz, dp, up]
    L.solarosc_eval_llh_sharded_device.argtypes = [H, C.c_void_p, sz, sz, C.c_void_p, C.c_void_p, C.c_void_p]
    L.solarosc_eval_asymmetry.argtypes = [H, C.c_int, dp, sz, dp, dp]
    L.solarosc_eval_llh_multi.argtypes = [C.POINTER(H), sz, dp, sz, sz, dp, C.POINTER(C.c_uint32)]
    L.solarosc_eval_derived.argtypes = [H, C.c_int, dp, sz] + [dp] * 11
    ip, u8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    L.solarosc_setup_bin_index.argtypes = [dp, sz, dp, sz, ip]
    L.solarosc_setup_bin_index_device.argtypes = [H, dp, sz, dp, sz, ip]
    L.solarosc_setup_histogram.argtypes = [dp, dp, u8, sz, C.c_int32, C.c_double, C.c_double, dp, C.c_int32, dp]
    L.solarosc_setup_response_matrix.argtypes = [dp, dp, sz, C.c_int32, C.c_double, C.c_double, dp, C.c_int32, C.c_double,
                                                 C.c_double, dp, dp, ip, ip]
    L.solarosc_setup_earth_paths.argtypes = [dp, dp, sz, dp, sz, C.c_int32, ip, ip, dp, sz, dp, ip]
    L.solarosc_eval_oscprobs.argtypes = [H, dp, sz, dp, dp, dp, dp]
    L.solarosc_eval_p1e.argtypes = [H, dp, sz, dp]
    L.solarosc_eval_llh_device.argtypes = [H, C.c_void_p, sz, sz, C.c_void_p, C.c_void_p, C.c_void_p]
    L.solarosc_poisson_llh.argtypes = [H, dp, dp, sz, sz, dp, up]
    L.solarosc_set_timing.argtypes = [H, C.c_int]
    L.solarosc_last_timing.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.solarosc_last_fused.argtypes = [H]
    L.solarosc_measure_fp64_peak.argtypes = [H, dp]
    L.solarosc_batcher_create.argtypes = [H, C.c_int, C.c_int, C.c_int]
    L.solarosc_batcher_create.restype = C.c_void_p
    L.solarosc_batcher_create_with.argtypes = [BATCH_FN, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.solarosc_batcher_create_with.restype = C.c_void_p
    L.solarosc_batcher_eval.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, up]
    L.solarosc_batcher_leave.argtypes = [C.c_void_p, C.c_int]
    L.solarosc_batcher_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.solarosc_batcher_destroy.argtypes = [C.c_void_p]
    L.solarosc_batcher_destroy.restype = None
    _lib = L
    return L
