#!/usr/bin/env python
"""N ranks (torchrun, one per GPU): the in-library NCCL gather (solarosc_eval_llh_sharded) returns on every rank what a
single-GPU evaluation of the whole batch returns -- uneven slices, grouped and ungrouped, flags included.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from bench import load_tables  # noqa: E402
from conftest import draw_params  # noqa: E402
from dune_solaroscfitter_b200.engine import SolarOscEngine  # noqa: E402
from oracle import oracle_np  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = load_tables()
ok = True
with SolarOscEngine(t, data=oracle_np.asimov_data(t), device=local) as e:
    e.comm_init_from_torch()
    for n_groups, group in ((1001, 1), (world * 500 + 1, 27), (3, 1), (70001, 1)):
        n = n_groups * group
        p = draw_params(t, n, seed=7)
        p[n // 2, 9] = 2.5                       # an out-of-bounds point: lnL = -Inf, flag 1
        lo, hi = e.shard_bounds(n, group)
        got, fl = e.eval_llh_sharded(p[lo:hi], n, group=group, return_flags=True)
        ref, rfl = e.eval_llh(p, group=group, return_flags=True)
        same = bool(np.array_equal(got, ref) and np.array_equal(fl, rfl))
        ok = ok and same
        print(f"rank {rank}/{world}: n = {n}, group = {group}, rows [{lo}, {hi}): gathered == direct: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_CHECK_OK" if int(flag.item()) == 1 else "SHARDED_CHECK_FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
