"""Rendezvous batcher (solarosc_batcher_*): N single-point callers -> one batched evaluation.  Host-side logic
only, driven through the C ABI with a CPU stand-in evaluator, so it runs without a GPU.  The reference has no
counterpart (its chains each call the CPU likelihood, mcmc.jl:96-110); the contract tested here is the one
include/solarosc_b200.h states."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

from dune_solaroscfitter_b200 import _lib

P = 5


@pytest.fixture(scope="module")
def L():
    return _lib.load()


def make_eval(log):
    def fn(ctx, params, n, lnl, grad, flags):
        p = np.ctypeslib.as_array(params, shape=(n, P))
        log.append(int(n))
        for i in range(n):
            lnl[i] = -float(np.sum(p[i] ** 2))
            flags[i] = 0
            if grad:
                for k in range(P):
                    grad[i * P + k] = -2.0 * p[i, k]
        return 0
    return _lib.BATCH_FN(fn)


def run_chain(L, b, slot, steps, out, leave=True):
    dp = C.POINTER(C.c_double)
    for s in range(steps):
        p = np.full(P, slot + 0.01 * s)
        lnl, fl = C.c_double(), C.c_uint32()
        g = np.empty(P)
        st = L.solarosc_batcher_eval(b, slot, p.ctypes.data_as(dp), C.byref(lnl), g.ctypes.data_as(dp), C.byref(fl))
        out.append((slot, s, st, lnl.value, g.copy(), p.copy()))
    if leave:
        assert L.solarosc_batcher_leave(b, slot) == 0


def test_full_batches_in_lockstep(L):
    log, out, n, steps = [], [], 8, 40
    cb = make_eval(log)
    b = L.solarosc_batcher_create_with(cb, None, P, n, 1, 0)
    assert b
    th = [threading.Thread(target=run_chain, args=(L, b, s, steps, out, False)) for s in range(n)]
    [t.start() for t in th]
    [t.join(60) for t in th]
    assert not any(t.is_alive() for t in th)
    assert log == [n] * steps                       # every launch carried all participants
    assert len(out) == n * steps
    for slot, s, st, lnl, g, p in out:
        assert st == 0 and lnl == -np.sum(p ** 2) and np.array_equal(g, -2 * p)
    nb, npnt = C.c_ulonglong(), C.c_ulonglong()
    L.solarosc_batcher_stats(b, C.byref(nb), C.byref(npnt))
    assert (nb.value, npnt.value) == (steps, n * steps)
    L.solarosc_batcher_destroy(b)


def test_chains_of_different_length_leave(L):
    log, out, n = [], [], 6
    cb = make_eval(log)
    b = L.solarosc_batcher_create_with(cb, None, P, n, 0, 0)
    steps = [5, 9, 13, 13, 2, 7]
    th = [threading.Thread(target=run_chain, args=(L, b, s, steps[s], out)) for s in range(n)]
    [t.start() for t in th]
    [t.join(60) for t in th]
    assert not any(t.is_alive() for t in th)        # nobody is left waiting for a chain that has finished
    assert sum(log) == sum(steps) and len(out) == sum(steps)
    assert max(log) <= n and len(log) >= max(steps)
    for slot, s, st, lnl, g, p in out:
        assert st == 0 and lnl == -np.sum(p ** 2)
    L.solarosc_batcher_destroy(b)


def test_patience_runs_out(L):
    log, out = [], []
    cb = make_eval(log)
    b = L.solarosc_batcher_create_with(cb, None, P, 3, 0, 50)      # three slots, only one ever calls
    t0 = time.time()
    run_chain(L, b, 1, 2, out, leave=False)
    assert 0.09 < time.time() - t0 < 5.0
    assert log == [1, 1] and [o[2] for o in out] == [0, 0]
    L.solarosc_batcher_destroy(b)


def test_argument_errors(L):
    cb = make_eval([])
    assert not L.solarosc_batcher_create_with(cb, None, 0, 4, 0, 0)
    b = L.solarosc_batcher_create_with(cb, None, P, 2, 0, 0)
    lnl = C.c_double()
    p = np.zeros(P)
    assert L.solarosc_batcher_eval(b, 5, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(lnl), None, None) == 1
    assert L.solarosc_batcher_leave(b, 0) == 0 and L.solarosc_batcher_leave(b, 0) == 1
    assert L.solarosc_batcher_eval(b, 0, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(lnl), None, None) == 1
    L.solarosc_batcher_destroy(b)
