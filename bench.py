#!/usr/bin/env python
"""bench.py -- full ES+CC day/night -2lnL evaluations per second (BASELINE.json's metric).

Workload (BASELINE.json configs[1]): the reference's `RunMode: LLH`
2-D scan of sin^2(theta12) x dm^2_21 (llhScan.jl:35-37,84-116) on a G x G grid (G = 512), every
cell evaluated at n_sys x 3 parameter points (-FWHM, +FWHM, centre of each of the 9 dominant
nuisances, llhScan.jl:48-70,94-108): 512 x 512 x 27 = 7.08 M parameter points per scan.  One
"step" is one pass over that batch; every point runs all five stages (no sharing of the
oscillation stage between the points of a cell -- that variant is reported under "grouped").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU; each rank scans the same (theta12, dm21) grid at its
own sin^2(theta13) value (weak scaling), and the per-point lnL vectors are all-gathered with NCCL
inside the timed region.  Rank 0 prints one JSON line.
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


def cpu_sample(t, params, seconds_budget: float = 12.0):
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
    return len(sample) / dt, cores, f"{len(sample)} evenly spaced points of the {len(params)}-point scan, {dt:.1f} s", lnl, sample


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is Julia
    and no Julia exists in this image (nor does its un-vendored NuFast-Earth dependency), so the arm
    times the C restatement (oracle/solarosc_oracle.c), OpenMP over parameter points."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t = load_tables()
    params, n_cell = scan_params(t, args.grid, float(t["truth"][1]))
    vals, sample_desc, cores = [], "", 1
    for it in range(args.warmup + args.steps):
        v, cores, sample_desc, _, _ = cpu_sample(t, params, seconds_budget=max(2.0, 60.0 / (args.warmup + args.steps)))
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * len(params) / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"LLH 2-D scan sin2_th12 x dm2_21, {args.grid}x{args.grid} cells x {n_cell} points (C2)",
                   "points_per_step": len(params), "n_params": params.shape[1]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_desc + " per step; C restatement of the Julia path (no Julia in the image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
    # weak scaling: rank r scans the same grid at its own sin2_th13 (llhScan.jl scans three such planes)
    s13 = float(t["truth"][1]) * (1.0 + 0.05 * rank)
    params, n_cell = scan_params(t, args.grid, s13)
    n, P = params.shape
    eng = SolarOscEngine(t, data=data, device=local)
    eng.set_timing(True)

    h_params = torch.from_numpy(params).pin_memory()
    d_params = h_params.cuda(non_blocking=False)
    d_lnl = torch.empty(n, dtype=torch.float64, device="cuda")
    d_flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    gathered = torch.empty(world * n, dtype=torch.float64, device="cuda") if world > 1 else None
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
        eng.eval_llh_device(d_params.data_ptr(), n, d_lnl.data_ptr(), d_flags.data_ptr(), group=group,
                            stream_ptr=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_lnl)

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
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- grouped variant: the 27 points of a cell share one oscillation stage --------------------------
    for _ in range(2):
        step_device(group=n_cell)
    ms_grouped = timed(lambda: step_device(group=n_cell), args.steps)
    grouped_value = world * n * args.steps / (ms_grouped * 1e-3)

    # ---- end to end through the host-buffer C ABI ("e2e") ---------------------------------------------
    h_np = h_params.numpy()
    out_np = h_lnl.numpy()

    def step_e2e():
        eng.eval_llh(h_np, out=out_np)         # H2D of the step's inputs + kernels + D2H of lnL, all inside
        if world > 1:
            dist.barrier()

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
    e2e_value = world * n * args.steps / float(dt.item())

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

    # ---- parity gate on this run's own output ------------------------------------------------------------
    step_device()                              # (the grouped pass left its own summation order in d_lnl)
    torch.cuda.synchronize()
    lnl_dev = d_lnl.cpu().numpy()
    same = bool(np.array_equal(lnl_dev, out_np))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_llh_kernel_traffic.json")))
        traffic = tr["dram_bytes_per_point"] * pts_per_launch
    except Exception:
        pass
    osc_s = ms_osc * 1e-3 / args.steps         # average duration of the first kernel's launches of one step
    rates_s = ms_rates * 1e-3 / args.steps
    if fused:   # one kernel does K1..K5
        ach = n * (FLOP_OSC + FLOP_RATES) / osc_s * 1e-12
        roofline = {"bound": "fp64", "kernel": "llh_kernel (fused K1-K5)", "achieved": ach, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic,
                    "traffic_note": "DRAM bytes per launch: ncu dram__bytes_read+write per point (profiles/r01_llh_kernel_traffic.json) "
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
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"LLH 2-D scan sin2_th12 x dm2_21, {args.grid}x{args.grid} cells x {n_cell} points (C2), "
                               "fastFit, ES+CC, 50 zenith x 80 energies x 6 Earth layers",
                   "points_per_step_per_gpu": n, "n_params": P,
                   "l2": f"inputs larger than L2 ({n * P * 8 / 1e6:.0f} MB of parameters per step per GPU)",
                   "parallelism": f"{world} x independent parameter slices + NCCL all-gather of lnL" if world > 1 else "1 GPU"},
        "roofline": roofline, "roofline_rates": roofline_rates,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * P * 8), "d2h_bytes_per_step": int(n * 12)},
        "grouped": {"value": grouped_value, "unit": UNIT, "group": n_cell,
                    "note": "oscillation stage shared by the points of a scan cell (precomputed_osc); not the headline"},
        "value_and_gradient": {"value": grad_value, "unit": "lnL + gradient evaluations/s", "batch": 4096,
                               "note": "solarosc_eval_llh_grad from host buffers on one GPU (HMC entry, C3); every evaluation "
                                       "runs 1 + 2 x 9 oscillation stages; not the headline"},
        "inclusive_angular": {"value": branch_value, "unit": UNIT, "batch": 32768,
                              "note": "VD-style analysis (inclusive CC + angular ES, 25 cos-scatter bins: 25 x the Poisson terms) through "
                                      "solarosc_eval_llh from host buffers (osc_kernel + gen_kernel); not the headline"},
        "gpu_launches": int(launches), "clocks": clocks,
        "e2e_equals_device_result": same,
    }
    if not args.no_cpu and world == 1:   # the CPU leg runs at N = 1 only
        v, cores, desc, lnl_cpu, sample = cpu_sample(t, params)
        gpu_on_sample = eng.eval_llh(sample)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line["parity_on_sample"] = {"max_abs_diff_m2lnl": float(np.max(np.abs(2 * gpu_on_sample - 2 * lnl_cpu))),
                                    "tolerance": 1e-6}
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
