"""Multi-rank sharding logic on CPU: world_size 2 over gloo, the oracle standing in for the engine."""
import os
import socket

import numpy as np
import pytest

from dune_solaroscfitter_b200.sharding import shard_bounds


def test_shard_bounds_cover_and_respect_groups():
    for n_groups, world, group in [(10, 3, 27), (8, 8, 1), (5, 8, 9), (1, 2, 27), (0, 4, 3)]:
        n = n_groups * group
        spans = [shard_bounds(n, world, r, group) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c
        assert all((b - a) % group == 0 for a, b in spans)
        sizes = [(b - a) // group for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 0, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, params, group, ret):
    import torch.distributed as dist
    from dune_solaroscfitter_b200.sharding import evaluate_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def fake_eval(p):           # a cheap pure function of the parameter row
            calls.append(len(p))
            return np.sin(p[:, 0] * 7.0) + p[:, 2] * 1e3

        out = evaluate_sharded(fake_eval, params, group=group)
        ret[rank] = (out, calls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_groups,group", [(7, 3), (2, 27), (1, 5)])
def test_two_rank_gather_matches_single_process(n_groups, group):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(3)
    params = rng.uniform(0.1, 0.5, (n_groups * group, 19))
    want = np.sin(params[:, 0] * 7.0) + params[:, 2] * 1e3
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, params, group, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        out, calls = ret[r]
        assert np.array_equal(out, want)            # every rank holds the full, ordered vector
        lo, hi = shard_bounds(len(params), 2, r, group)
        assert calls == ([hi - lo] if hi > lo else [])   # and evaluated only its own slice


def test_library_shard_bounds_match_the_python_ones():
    """solarosc_shard_bounds (the split the in-library NCCL gather uses) against sharding.shard_bounds, and the inverse
    map row -> rank that the unpadding kernel computes (csrc/solarosc_comm.hpp), restated here."""
    import ctypes as C
    from dune_solaroscfitter_b200 import _lib
    from dune_solaroscfitter_b200.sharding import shard_bounds
    L = _lib.load()
    for n_groups in (0, 1, 5, 8, 27, 1000, 262144):
        for group in (1, 27):
            for world in (1, 2, 3, 4, 8):
                n = n_groups * group
                owner = np.full(n, -1)
                for r in range(world):
                    lo, hi = C.c_size_t(), C.c_size_t()
                    assert L.solarosc_shard_bounds(n, group, r, world, C.byref(lo), C.byref(hi)) == 0
                    assert (lo.value, hi.value) == shard_bounds(n, world, r, group)
                    owner[lo.value:hi.value] = r
                if n == 0:
                    continue
                base, extra = divmod(n_groups, world)
                gi = np.arange(n) // group
                cut = extra * (base + 1)
                r = np.where(gi < cut, gi // (base + 1), extra + (gi - cut) // max(base, 1))
                assert np.array_equal(r, owner)
    lo, hi = C.c_size_t(), C.c_size_t()
    assert L.solarosc_shard_bounds(10, 3, 0, 2, C.byref(lo), C.byref(hi)) == 1      # n not a multiple of group
