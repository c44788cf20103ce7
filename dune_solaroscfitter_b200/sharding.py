"""Multi-GPU sharding of a parameter batch: one process per GPU, no data-path collective except the
gather of the per-point lnL vector.

Every parameter point is independent of every other (``total_llh`` has no cross-point state,
reference likelihood_main.jl:30-35), so the batch is split by contiguous rows, in units of ``group``
points so that the points of one LLH-scan cell stay on one GPU.  Tables are
replicated: each rank creates its own ``SolarOscEngine`` on its own device.  The only communication is
one all-gather of ``lnL`` (8 bytes per point) over NCCL/NVLink; under ``gloo`` the same code runs on
CPU tensors, which is how the logic is tested without GPUs.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_bounds(n_points: int, world: int, rank: int, group: int = 1) -> Tuple[int, int]:
    """Half-open row range of ``rank``: groups are dealt out as evenly as possible, earlier ranks first."""
    if n_points % group:
        raise ValueError("n_points must be a multiple of group")
    n_groups = n_points // group
    base, extra = divmod(n_groups, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * group, hi * group


def evaluate_sharded(eval_fn: Callable[[np.ndarray], np.ndarray], params: np.ndarray, group: int = 1,
                     device=None) -> np.ndarray:
    """Evaluate this rank's slice with ``eval_fn`` and all-gather the full lnL vector on every rank.

    ``eval_fn(params_slice) -> lnL_slice`` is normally ``SolarOscEngine.eval_llh``.  Falls back to a
    plain call when ``torch.distributed`` is not initialised (single process).
    """
    import torch
    import torch.distributed as dist

    n = len(params)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(eval_fn(params))
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n, world, rank, group)
    mine = np.asarray(eval_fn(params[lo:hi]), dtype=np.float64) if hi > lo else np.empty(0)
    width = max(shard_bounds(n, world, r, group)[1] - shard_bounds(n, world, r, group)[0] for r in range(world))
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=device)
    buf[: hi - lo] = torch.from_numpy(mine).to(buf.device)
    out = torch.empty(world * width, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, width)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n, world, r, group)
        parts.append(out[r, : b - a])
    return np.concatenate(parts)
