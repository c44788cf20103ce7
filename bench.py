#!/usr/bin/env python
"""bench.py -- full ES+CC day/night -2lnL evaluations per second (BASELINE.json's metric).

Workload (BASELINE.json configs[1]): the reference's `RunMode: LLH`
2-D scan of sin^2(theta12) x dm^2_21 (llhScan.jl:35-37,84-116) on a G x G grid (G = 512), every
cell evaluated at n_sys x 3 parameter points (-FWHM, +FWHM, centre of each of the 9 dominant
nuisances, llhScan.jl:48-70,94-108): 512 x 512 x 27 = 7.08 M parameter points per scan.  One
"step" is one pass over that batch; every point runs all five stages (no sharing of the
oscillation stage between the points of a cell -- that variant is reported under "grouped").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl reference] [--scaling weak|strong]

N > 1 is launched by torchrun, one rank per GPU.  --scaling weak (default, what the driver runs): each rank scans the
same (theta12, dm21) grid at its own sin^2(theta13) value; --scaling strong: ONE scan is split over the ranks by
solarosc_shard_bounds.  Either way the per-point lnL and flag vectors are all-gathered device-to-device by the
library's own NCCL communicator (solarosc_eval_llh_sharded_device) inside the timed region; torch.distributed only
bootstraps it and provides the barriers.  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ES+CC day/night -2lnL evaluations/sec"
UNIT = "evaluations/s"
# Algorithmic FP64 work per evaluation (fast mode, 1 FMA = 2 flop, sincos/log/sqrt
# not counted): K1 MSW 0.005 + K2 Earth 2.57 + night assembly 0.03 MFLOP in osc_kernel;
# K3 0.04 + K4 0.367 + K5 0.04 MFLOP in rates_kernel.
FLOP_OSC = 2.605e6
FLOP_RATES = 0.447e6
SINCOS_PER_EVAL = 46560       # the reference's three complex exponentials per layer step
# Stage-separated HBM bytes per evaluation of the rate kernel: oscillation probabilities read once
# (2 x 51 x 40 doubles), parameters in, lnL + flag out.
BYTES_RATES = 2 * 51 * 40 * 8 + 19 * 8 + 12


def workload_config(grid: int, n_cell: int, n: int, P: int, world: int, scaling: str):
    """`config` of the JSON line: the same object in both arms (the driver compares them)."""
    return {"workload": f"LLH 2-D scan sin2_th12 x dm2_21, {grid}x{grid} cells x {n_cell} points (C2), "
                        "fastFit, ES+CC, 50 zenith x 80 energies x 6 Earth layers",
            "points_per_step_per_gpu": n, "n_params": P,
            "l2": f"inputs larger than L2 ({n * P * 8 / 1e6:.0f} MB of parameters per step per GPU)",
            "parallelism": (f"{world} x {'independent parameter slices' if scaling == 'weak' else 'slices of one scan'} + "
                            "NCCL all-gather of lnL inside the library") if world > 1 else "1 GPU"}


def load_tables():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "tables_escc.npz")))


def scan_params(t, grid: int, sin2_th13: float) -> np.ndarray:
    """Parameter points of one 2-D LLH scan in llhScan.jl's loop order (cell-major, 27 per cell)."""
    truth = np.asarray(t["truth"], dtype=np.float64)
    nl = len(t["earth_lookup"])
    v12 = np.linspace(0.2, 0.4, grid)          # llhScan.jl:35,40
    vdm = np.linspace(4e-5, 1.5e-4, grid)      # llhScan.jl:37,42
    k = math.sqrt(2 * math.log(2))
    # nuisance column -> FWHM shift (llhScan.jl:48-70); prior widths from config.yaml:38-52
    shifts = [(3, k * 0.28e6), (5, k * 0.1), (6, k * 0.1), (7, k * 0.05)]
    shifts += [(c, k * 0.02 * truth[c]) for c in range(8 + nl, len(truth))]
    cell = np.tile(truth, (3 * len(shifts), 1))
    for s, (col, d) in enumerate(shifts):
        cell[3 * s + 0, col] -= d
        cell[3 * s + 1, col] += d
    cell[:, 1] = sin2_th13
    n_cell = len(cell)
    p = np.empty((grid, grid, n_cell, len(truth)))
    p[:] = cell
    p[:, :, :, 0] = v12[:, None, None]
    p[:, :, :, 2] = vdm[None, :, None]
    return p.reshape(-1, len(truth)), n_cell


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as exc:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_sample(t, params, seconds_budget: float = 12.0, single_thread: bool = False):
    """Time the C oracle (all host threads) on a bounded, evenly spaced sample of the workload."""
    from oracle import oracle_c, oracle_np
    co = oracle_c.COracle(t, oracle_np.asimov_data(t))
    # every host core the process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS, not OpenMP)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    probe = params[:: max(1, len(params) // (4 * cores))][: 4 * cores]
    t0 = time.perf_counter()
    co.eval_llh(probe, n_threads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(max(cores, min(len(params), rate * seconds_budget)))
    sample = np.ascontiguousarray(params[:: max(1, len(params) // n)][:n])
    t0 = time.perf_counter()
    lnl, _ = co.eval_llh(sample, n_threads=cores)
    dt = time.perf_counter() - t0
    one = None
    if single_thread:   # single-thread ms per evaluation
        k = sample[:: max(1, len(sample) // 64)][:64]
        t0 = time.perf_counter()
        co.eval_llh(np.ascontiguousarray(k), n_threads=1)
        one = 1e3 * (time.perf_counter() - t0) / len(k)
    return dict(value=len(sample) / dt, cores=cores, n=len(sample), seconds=dt, lnl=lnl, sample=sample, single_thread_ms_per_eval=one,
                desc=f"{len(sample)} evenly spaced points of the {len(params)}-point scan, {dt:.1f} s")


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is Julia
    and no Julia exists in this image (nor does its un-vendored NuFast-Earth dependency), so the arm
    times the C restatement (oracle/solarosc_oracle.c), OpenMP over parameter points.  One step = one bounded,
    evenly spaced sample of the scan (stated in `reference_sample`); ms_per_step is the time that sample took."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t = load_tables()
    params, n_cell = scan_params(t, args.grid, float(t["truth"][1]))
    vals, ms, last = [], [], None
    for it in range(args.warmup + args.steps):
        last = cpu_sample(t, params, seconds_budget=max(2.0, 60.0 / (args.warmup + args.steps)), single_thread=(it == 0))
        one = last["single_thread_ms_per_eval"] if it == 0 else one
        if it >= args.warmup:
            vals.append(last["value"]); ms.append(1e3 * last["seconds"])
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.grid, n_cell, len(params), params.shape[1], 1, args.scaling),
        "reference_sample": {"points_per_step": last["n"], "of": len(params),
                             "note": "each step evaluates this evenly spaced sample of the scan on all host cores; value = points / time, "
                                     "the whole scan would take ms_per_step x of / points_per_step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": last["desc"] + " per step; C restatement of the Julia path (no Julia in the image)",
                         "single_thread_ms_per_eval": one},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def extra_legs(eng, t, params, local):
    """The other BASELINE.json configurations, measured in the same run (N = 1): small, bounded legs."""
    import torch
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import draw_params
    out = {}

    def best(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    # C1 / C3: latency of small batches from host buffers (one captured CUDA graph per call), MCMC-sized batches
    truth = np.asarray(t["truth"], dtype=float)[None, :]
    p10, p4096 = draw_params(t, 10, seed=3), draw_params(t, 4096, seed=4)
    out["c1_single_point"] = {"us_per_call": 1e6 * best(lambda: eng.eval_llh(truth), 20), "note": "one point at truth from host buffers (C1)"}
    tv, tg = best(lambda: eng.eval_llh(p4096)), best(lambda: eng.eval_llh_grad(p4096))
    out["c3_mcmc"] = {"b10_us_per_step": 1e6 * best(lambda: eng.eval_llh(p10), 20), "b4096_ms_per_step": 1e3 * tv,
                      "b4096_value_per_s": 4096 / tv, "b10_grad_us_per_step": 1e6 * best(lambda: eng.eval_llh_grad(p10), 20),
                      "b4096_grad_ms_per_step": 1e3 * tg, "b4096_value_and_gradient_per_s": 4096 / tg,
                      "note": "independent prior draws per step from host buffers (C3); gradient = exact adjoint sweep"}
    # C5: derived-mode replay of a stored chain, block sums on the device
    n5 = 150_000
    p5 = np.ascontiguousarray(np.tile(draw_params(t, 15_000, seed=5), (10, 1)))
    lref = {k: v[0] for k, v in eng.eval_perbin(truth).items()}
    w5 = np.ones(n5)
    ta, td = best(lambda: eng.eval_asymmetry(p5), 2), best(lambda: eng.eval_derived(p5, w5, lref), 2)
    out["derived_c5"] = {"samples": n5, "asymmetry_samples_per_s": n5 / ta, "block_sums_samples_per_s": n5 / td,
                         "note": "D/N asymmetries and per-bin weighted moments of a 1.5e5-sample chain, reduced on the device (C5)"}
    # fastFit: false (production-region integral)
    ps = np.ascontiguousarray(params[:: max(1, len(params) // 110592)][:110592])
    with SolarOscEngine(t, data=oracle_np.asimov_data(t), device=local, fast=False) as es:
        out["slow_mode"] = {"value": len(ps) / best(lambda: es.eval_llh(ps), 2), "unit": UNIT, "batch": len(ps),
                            "note": "fastFit: false, 1000 solar radii per energy"}
    # C4: night-only CC likelihood with earth-norm draws on dense zenith x energy grids (utils/checkOscGridConvergence.jl:36-46)
    c4 = {}
    for name, nb in (("tables_dense200.npz", 2048), ("tables_dense1000.npz", 96)):
        td_ = dict(np.load(os.path.join(ROOT, "tests", "golden", name)))
        with SolarOscEngine(td_, device=local, es_mode=False, night_only=True) as ed:
            tr = np.asarray(td_["truth"], dtype=float)
            r0 = ed.eval_rates(tr[None, :])
            ed.set_data({k: r0[k][0] for k in ("ES_day", "ES_night", "CC_day", "CC_night")})
            pd_ = np.tile(tr, (nb, 1))
            pd_[:, 8:14] = np.clip(np.random.default_rng(3).normal(1.0, 0.12, (nb, 6)), 0.0, 2.0)
            pd_[:, 0], pd_[:, 2] = 0.4, 2e-4                                  # worst-case oscillation point of the sweep
            dt = best(lambda: ed.eval_llh(pd_), 2)
            steps = int(np.diff(td_["seg_start"]).sum()) * len(td_["e_calc"])
            flop = FLOP_OSC * steps / (80 * 194)                              # K2 scaled by steps per evaluation
            c4[name.replace("tables_", "").replace(".npz", "")] = {
                "value": nb / dt, "unit": UNIT, "batch": nb, "grid": f"{len(td_['exposure_weights'])} zenith x {len(td_['e_calc'])} energies",
                "layer_steps_per_eval": steps, "algorithmic_flop_per_eval": flop, "achieved_tflops": nb * flop / dt * 1e-12}
    out["c4_dense"] = c4
    out["c4_dense"]["note"] = ("night-only CC likelihood (solarosc_config.night_only), six earth norms ~ truncN(1, 0.12), host buffers: "
                               "osc_kernel with the energy axis split over CTAs + row-chunked fold; FP64 fraction = achieved_tflops / roofline.peak")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1/C3/C4/C5/slow-mode legs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dune_solaroscfitter_b200.engine import SolarOscEngine
    from oracle import oracle_np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t = load_tables()
    data = oracle_np.asimov_data(t)            # Asimov data = rates at truth (setup.jl:160-166); fixed input table
    strong = args.scaling == "strong" and world > 1
    eng = SolarOscEngine(t, data=data, device=local)
    eng.set_timing(True)
    if world > 1:
        eng.comm_init_from_torch()             # the library's own NCCL communicator (solarosc_comm_init)
    if strong:     # ONE scan, split over the ranks in whole cells
        full, n_cell = scan_params(t, args.grid, float(t["truth"][1]))
        n_total = len(full)
        lo, hi = eng.shard_bounds(n_total, n_cell)
        params = np.ascontiguousarray(full[lo:hi])
        del full
    else:          # weak scaling: rank r scans the same grid at its own sin2_th13 (llhScan.jl scans three such planes)
        s13 = float(t["truth"][1]) * (1.0 + 0.05 * rank)
        params, n_cell = scan_params(t, args.grid, s13)
        n_total = world * len(params)
    n, P = params.shape

    h_params = torch.from_numpy(params).pin_memory()
    d_params = h_params.cuda(non_blocking=False)
    d_lnl = torch.empty(n, dtype=torch.float64, device="cuda")
    d_flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    gathered = torch.empty(n_total, dtype=torch.float64, device="cuda") if world > 1 else None
    gathered_fl = torch.zeros(n_total, dtype=torch.int32, device="cuda") if world > 1 else None
    h_lnl = torch.empty(n, dtype=torch.float64).pin_memory()
    # a real (non-default) stream: its handle is what the library launches on and what the events time
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(group=1):
        if world > 1:   # this rank's slice, then the library's NCCL all-gather of lnL and flags, device to device
            eng.eval_llh_sharded_device(d_params.data_ptr(), n_total, gathered.data_ptr(), gathered_fl.data_ptr(), group=group,
                                        stream_ptr=stream.cuda_stream)
        else:
            eng.eval_llh_device(d_params.data_ptr(), n, d_lnl.data_ptr(), d_flags.data_ptr(), group=group,
                                stream_ptr=stream.cuda_stream)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput ("value") ------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms_osc = ms_rates = 0.0
    launches = 0
    fused = False
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        tm = eng.last_timing()                 # waits for this step's kernels; per-kernel CUDA events
        ms_osc += tm["ms_osc"]; ms_rates += tm["ms_rates"]; launches += tm["launches"]; fused = tm["fused"]
    e1.record(stream)
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    clocks = sampler.stop()
    value = n_total * args.steps / (ms_total * 1e-3)

    # ---- grouped variant: the 27 points of a cell share one oscillation stage --------------------------
    for _ in range(2):
        step_device(group=n_cell)
    ms_grouped = timed(lambda: step_device(group=n_cell), args.steps)
    grouped_value = n_total * args.steps / (ms_grouped * 1e-3)

    # ---- end to end through the host-buffer C ABI ("e2e") ---------------------------------------------
    h_np = h_params.numpy()
    out_np = h_lnl.numpy()

    # every rank receives the results of ALL ranks: pinned, caller-owned arrays handed to the library
    out_all = torch.empty(n_total, dtype=torch.float64).pin_memory().numpy() if world > 1 else None

    def step_e2e():
        if world > 1:   # host rows in, all ranks' lnL out: H2D + kernels + NCCL gather + D2H inside the call
            eng.eval_llh_sharded(h_np, n_total, out=out_all)   # lnL of all ranks; a negative-rate error would come back as the status
        else:
            eng.eval_llh(h_np, out=out_np)     # H2D of the step's inputs + kernels + D2H of lnL, all inside

    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = n_total * args.steps / float(dt.item())

    # ---- value + gradient (C3, HMC): 4096 chain points per call through the host-buffer C ABI ---------------
    grad_value = None
    if rank == 0:
        gp = np.ascontiguousarray(params[:: max(1, n // 4096)][:4096])
        eng.eval_llh_grad(gp)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.eval_llh_grad(gp)
        grad_value = 3 * len(gp) / (time.perf_counter() - t0)

    # ---- the other analysis branches (VD block of config.yaml: inclusive + angular ES, 25 cos-scatter bins) ----------
    branch_value = None
    if rank == 0:
        ext = dict(np.load(os.path.join(ROOT, "tests", "golden", "tables_ext.npz")))
        tv = dict(t)
        tv.update({k: ext[k] for k in ("n_cos", "resp_cc_incl", "eff_cc_incl", "resp_es_angular", "resp_bg_angular", "angular_mask")})
        tv["eff_es_nue"], tv["eff_es_nuother"] = ext["eff_es_nue_boosted"], ext["eff_es_nuother_boosted"]
        tv["inclusive"], tv["angular"] = np.int32(1), np.int32(1)
        bp = np.ascontiguousarray(params[:: max(1, n // 32768)][:32768])
        with SolarOscEngine(tv, data=oracle_np.asimov_data(tv), device=local) as ev:
            ev.eval_llh(bp)
            t0 = time.perf_counter()
            for _ in range(2):
                ev.eval_llh(bp)
            branch_value = 2 * len(bp) / (time.perf_counter() - t0)

    # ---- strong scaling of an MCMC step (C3): B = 4096 chain points split over the ranks ------------------------
    c3_strong = None
    if strong:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import draw_params
        pc = draw_params(t, 4096, seed=4)
        clo, chi = eng.shard_bounds(4096, 1)
        d_pc = torch.from_numpy(np.ascontiguousarray(pc[clo:chi])).cuda()
        d_c3 = torch.empty(4096, dtype=torch.float64, device="cuda")
        fn = lambda: eng.eval_llh_sharded_device(d_pc.data_ptr(), 4096, d_c3.data_ptr(), 0, stream_ptr=stream.cuda_stream)  # noqa: E731
        for _ in range(5):
            fn()
        ms_c3 = timed(fn, 50) / 50
        # value + gradient: every rank differentiates its slice (host buffers), gradient rows all-gathered by torch (NCCL)
        gl = torch.zeros((4096 // world + 1, P + 1), dtype=torch.float64, device="cuda")
        ga = torch.zeros((world * (4096 // world + 1), P + 1), dtype=torch.float64, device="cuda")

        def grad_step():
            l, g = eng.eval_llh_grad(pc[clo:chi])
            gl[: chi - clo, 0] = torch.from_numpy(l).cuda(); gl[: chi - clo, 1:] = torch.from_numpy(g).cuda()
            dist.all_gather_into_tensor(ga, gl)
        for _ in range(3):
            grad_step()
        barrier(); t0 = time.perf_counter()
        for _ in range(10):
            grad_step()
        torch.cuda.synchronize(); dtg = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device="cuda")
        dist.all_reduce(dtg, op=dist.ReduceOp.MAX)
        c3_strong = {"batch": 4096, "value_ms_per_step": ms_c3, "value_per_s": 4096 / (ms_c3 * 1e-3),
                     "value_and_gradient_ms_per_step": 1e3 * float(dtg.item()), "value_and_gradient_per_s": 4096 / float(dtg.item()),
                     "note": f"{4096 // world} chain points per GPU: the value step is bound by kernel launch + ncclAllGather latency and by "
                             "filling 148 SMs with a few hundred points; the gradient step by its three kernels on the same few points"}

    # ---- parity gate on this run's own output ------------------------------------------------------------
    step_device()                              # (the grouped pass left its own summation order in d_lnl)
    torch.cuda.synchronize()
    if world > 1:
        lo_r, hi_r = (eng.shard_bounds(n_total, n_cell) if strong else (rank * n, (rank + 1) * n))
        lnl_dev = gathered[lo_r:hi_r].cpu().numpy()
        same = bool(np.array_equal(gathered.cpu().numpy(), out_all))
    else:
        lnl_dev = d_lnl.cpu().numpy()
        same = bool(np.array_equal(lnl_dev, out_np))

    def shutdown():
        """Every rank leaves together: the library's communicator is released collectively (solarosc_comm_destroy),
        then torch's.  A watchdog ends the process if a teardown call blocks -- the JSON line is out by then."""
        sys.stdout.flush()
        dog = threading.Timer(45.0, lambda: os._exit(0))
        dog.daemon = True
        dog.start()
        if world > 1:
            dist.barrier()
            eng.comm_destroy()
        eng.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        dog.cancel()

    if rank != 0:
        return shutdown()

    fp64_peak = eng.measure_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic of the dominant kernel from the committed ncu capture (bytes per point x points per launch)
    traffic, pts_per_launch = None, n * args.steps / max(1, launches)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_llh_kernel_traffic.json")))
        traffic = tr["dram_bytes_per_point"] * pts_per_launch
    except Exception:
        pass
    osc_s = ms_osc * 1e-3 / args.steps         # average duration of the first kernel's launches of one step
    rates_s = ms_rates * 1e-3 / args.steps
    if fused:   # one kernel does K1..K5
        ach = n * (FLOP_OSC + FLOP_RATES) / osc_s * 1e-12
        roofline = {"bound": "fp64", "kernel": "llh_kernel (fused K1-K5)", "achieved": ach, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic,
                    "traffic_note": "DRAM bytes per launch: ncu dram__bytes_read+write per point (profiles/r02_llh_kernel_traffic.json) "
                                    f"x {pts_per_launch:.0f} points per launch; algorithmic bytes are 164 per point",
                    "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "ms_per_step": ms_osc / args.steps, "share_of_step": 1.0,
                    "sincos_per_s": n * SINCOS_PER_EVAL / osc_s,
                    "algorithmic_flop_per_eval": FLOP_OSC + FLOP_RATES}
        roofline_rates = None
    else:
        ach = n * FLOP_OSC / osc_s * 1e-12
        roofline = {"bound": "fp64", "kernel": "osc_kernel", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak, "traffic": None,
                    "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "ms_per_step": ms_osc / args.steps, "share_of_step": ms_osc / (ms_osc + ms_rates),
                    "sincos_per_s": n * SINCOS_PER_EVAL / osc_s, "algorithmic_flop_per_eval": FLOP_OSC}
        ach_b = n * BYTES_RATES / rates_s * 1e-9
        roofline_rates = {"bound": "hbm", "kernel": "rates_kernel", "achieved": ach_b, "peak": hbm_peak, "unit": "GB/s",
                          "frac": ach_b / hbm_peak, "traffic": None,
                          "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                          "ms_per_step": ms_rates / args.steps, "fp64_tflops": n * FLOP_RATES / rates_s * 1e-12}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.grid, n_cell, n if world > 1 else len(params), P, world, args.scaling),
        "roofline": roofline, "roofline_rates": roofline_rates,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * P * 8), "d2h_bytes_per_step": int(n_total * 8 if world > 1 else n * 12)},
        "grouped": {"value": grouped_value, "unit": UNIT, "group": n_cell,
                    "note": "oscillation stage shared by the points of a scan cell (precomputed_osc); not the headline"},
        "value_and_gradient": {"value": grad_value, "unit": "lnL + gradient evaluations/s", "batch": 4096,
                               "note": "solarosc_eval_llh_grad from host buffers on one GPU (HMC entry, C3); exact gradient: reverse-mode "
                                       "fold + adjoint sweep through the Earth propagation; not the headline"},
        "inclusive_angular": {"value": branch_value, "unit": UNIT, "batch": 32768,
                              "note": "VD-style analysis (inclusive CC + angular ES, 25 cos-scatter bins: 25 x the Poisson terms) through "
                                      "solarosc_eval_llh from host buffers (osc_kernel + gen_kernel); not the headline"},
        "gpu_launches": int(launches), "clocks": clocks,
        "e2e_equals_device_result": same,
    }
    if c3_strong is not None:
        line["c3_strong"] = c3_strong
    if world == 1 and not args.no_extras:
        line.update(extra_legs(eng, t, params, local))
    if not args.no_cpu and world == 1:   # the CPU leg runs at N = 1 only
        cs = cpu_sample(t, params, single_thread=True)
        gpu_on_sample = eng.eval_llh(cs["sample"])
        line["cpu_baseline"] = {"value": cs["value"], "unit": UNIT, "cores": cs["cores"], "kind": "port", "sample": cs["desc"],
                                "single_thread_ms_per_eval": cs["single_thread_ms_per_eval"]}
        line["parity_on_sample"] = {"max_abs_diff_m2lnl": float(np.max(np.abs(2 * gpu_on_sample - 2 * cs["lnl"]))),
                                    "tolerance": 1e-6}
    print(json.dumps(line), flush=True)
    shutdown()


if __name__ == "__main__":
    main()
