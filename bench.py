#!/usr/bin/env python
"""bench.py -- headline benchmark: transient sweep points/s on the DFF Monte-Carlo workload.

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm on the box's host cores

One "step" = one full pass of the hot path over one batch: DC operating point + adaptive
transient (0 .. 600 ns, LTE-controlled trapezoidal) of B = 16 384 Monte-Carlo instances of the
30-FET D flip-flop per GPU (BASELINE.json configs[2]; GF180/BSIM4 are not in the reference tree,
so the same topology runs on BSIM-CMG 107 + ASAP7 cards -- SURVEY.md fact 5).

`value` is timed with inputs resident in HBM and results left in HBM; `e2e` goes through the
C-ABI calls with pinned host buffers (H2D of the parameter matrix and D2H of all waveforms inside
the timed region).  Under torchrun every rank runs its own block of sweep points (no data-path
collective, weak scaling) and rank 0 gathers the waveforms over NCCL.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DFF_NODESET = dict(q=0.0, q_neg=0.7, net0=0.0, net7=0.0, vdd=0.7, clkn=0.7, ncki=0.0, cki=0.7)
T0, T1, NSAVE = 0.0, 6e-7, 1801
POINTS_PER_GPU = 16384
# SURVEY.md 8(d) config 3, adaptive mode: reltol 1e-4, abstol 1e-6 V / 1e-12 A.  Newton tolerance = 0.1 x the LTE
# tolerance with the contraction-rate acceptance test (Sundials IDA, the reference's solver, uses 0.33 x and the same
# test); the CPU arm runs the same options.  value_rounds is the engine's chord-iteration schedule.
OPTS = dict(reltol=1e-4, vabstol=1e-6, iabstol=1e-12, nr_reltol=1e-5, nr_vabstol=1e-7, nr_iabstol=1e-13, nr_rate_test=1)
ENGINE_OPTS = dict(value_rounds=2)
WORKLOAD = ("dff30-bsimcmg107-asap7 monte-carlo transient, adaptive trap, reltol 1e-4 / 1e-6 V / 1e-12 A, Newton tol 0.1x LTE tol "
            "with rate test, 0..600ns, S=1801 (stand-in for GF180 DFF)")


def nodeset(fc):
    from cedarsim.jl_b200.flat import nodeset_vector
    return nodeset_vector(fc, DFF_NODESET)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                     "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag.is_set():
            line = proc.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class DevArray:
    """zero-copy view of engine-owned HBM for torch (NCCL gather)"""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def cpu_arm(points, threads):
    """The CPU restatement (oracle, kind 'port') on `points` sweep points with `threads` host threads."""
    from cedarsim.jl_b200 import circuits
    from oracle import orc
    fc, _ = circuits.dff(host=True)
    P = circuits.dff_mc_params(fc, points)
    orc.set_x0(nodeset(fc))
    ts = np.linspace(T0, T1, NSAVE)
    t = time.perf_counter()
    y, st, stats = orc.tran(fc, T0, T1, ts, params=P, opts=orc.default_options(**OPTS), nthreads=threads)
    el = time.perf_counter() - t
    orc.set_x0(None)
    return el, stats, int((st == 0).sum())


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    points = max(threads * 16, 64)   # ~10 s of CPU work per step on the box's host cores
    for _ in range(args.warmup):
        cpu_arm(points, threads)
    t = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        el, stats, ok = cpu_arm(points, threads)
        iters += stats["newton_iters"]
    el = time.perf_counter() - t
    value = args.steps * points / el
    print(json.dumps({
        "impl": "reference", "metric": "transient sweep points/s (DFF Monte-Carlo)", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_step": points},
        "newton_iters_per_s": iters / el,
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{points} of the {POINTS_PER_GPU} Monte-Carlo points per step, same tolerances and outputs"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--points", type=int, default=POINTS_PER_GPU, help="sweep points per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from cedarsim.jl_b200 import circuits, engine, models

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.points
    fc, ms = circuits.dff()
    circuit = engine.Circuit(fc, ms)
    plan = circuit.plan(B, device=local)
    # per-rank Monte-Carlo draws: rank r owns sweep points [r*B, (r+1)*B)
    P_all = circuits.dff_mc_params(fc, B * world)
    P = np.ascontiguousarray(P_all[:, rank * B:(rank + 1) * B])
    del P_all
    ts = np.linspace(T0, T1, NSAVE)
    opts = engine.default_options(**OPTS, **ENGINE_OPTS)
    plan.set_x0(nodeset(fc))
    plan.set_params(P)           # inputs resident in HBM before the timed region
    O = len(fc.outputs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather_buf = [torch.empty((O, NSAVE, B), dtype=torch.float64, device="cuda") for _ in range(world)] if (world > 1 and rank == 0) else None

    def step_resident():
        dy, ds, st = plan.tran_device(T0, T1, ts, opts)
        if world > 1:   # final waveforms to rank 0 over NVLink (NCCL); no collective on the solve path
            y = torch.as_tensor(DevArray(dy, (O, NSAVE, B)), device="cuda")
            dist.gather(y, gather_buf, dst=0)
        return st

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    tot = {"newton_iters": 0, "kernel_launches": 0, "eval_seconds": 0.0, "newton_seconds": 0.0, "solve_seconds": 0.0,
           "steps_accepted": 0, "steps_rejected": 0, "rounds": 0, "value_rounds": 0, "full_iters": 0, "evalv_seconds": 0.0,
           "newtonv_seconds": 0.0}
    for _ in range(args.steps):
        st = step_resident()
        for k in tot:
            tot[k] += st[k]
    barrier()
    el = time.perf_counter() - t0
    sampler.stop_flag.set()
    # ---- per-kernel timing for the roofline: the same step on a single-lane plan with CUDA events around every launch
    # (in the timed region above the plan's lanes run concurrently, so a kernel's launch duration there includes the
    # SMs it shares with other lanes' kernels; timed alone it is the figure the roofline peak is quoted for)
    os.environ["CB_EVAL_FORK"] = "0"   # one stream: the n- and p-FET eval kernels of a round run one after the other
    plan1 = circuit.plan(B, device=local, lanes=1)
    del os.environ["CB_EVAL_FORK"]
    plan1.set_x0(nodeset(fc))
    plan1.set_params(P)
    plan1.tran_device(T0, T1, ts, opts)
    plan1.set_timing(True)
    _, _, st1 = plan1.tran_device(T0, T1, ts, opts)
    plan1.close()
    for k in ("eval_seconds", "newton_seconds", "evalv_seconds", "newtonv_seconds", "solve_seconds", "full_iters", "value_rounds", "rounds",
              "newton_iters"):
        tot[k + "_1"] = st1[k]
    if world > 1:
        tmax = torch.tensor([el, tot["solve_seconds"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        el = float(tmax[0])
        sums = torch.tensor([tot["newton_iters"], tot["kernel_launches"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(sums)
        newton_all, launches_all = float(sums[0]), float(sums[1])
    else:
        newton_all, launches_all = float(tot["newton_iters"]), float(tot["kernel_launches"])
    value = args.steps * B * world / el

    # ---- end-to-end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    y_host = torch.empty((O, NSAVE, B), dtype=torch.float64, pin_memory=True).numpy()
    P_host = torch.from_numpy(P).pin_memory().numpy()
    plan.set_params(P_host); plan.tran(T0, T1, ts, opts, out=y_host)   # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.set_params(P_host)
        y_host, status, _ = plan.tran(T0, T1, ts, opts, out=y_host)
    barrier()
    el_e2e = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([el_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        el_e2e = float(tmax[0])
    ok_points = int((status == 0).sum())
    q_known = [float(np.abs(y_host[0, int(round(tt / T1 * (NSAVE - 1)))] - want).max())
               for tt, want in ((1.5e-7, 0.0), (2.5e-7, 0.0), (4.5e-7, 0.7), (5.5e-7, 0.7))]

    if rank == 0:
        # ---- roofline of the dominant kernel (device evaluation, FP64-pipe bound: SURVEY.md 8(d))
        flops_per_eval = float(np.mean([sum(cm.exec_ops) for cm in ms])) if all(getattr(cm, "exec_ops", None) for cm in ms) else None
        n_fets = len(fc.va_insts)
        fp64_peak = engine.measure_fp64_peak(local)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("k_eval_dram_bytes_per_launch")
        except OSError:
            pass
        ev, nw = tot["eval_seconds_1"], tot["newton_seconds_1"]
        solve1 = max(tot["solve_seconds_1"], 1e-30)
        # k_eval_* runs in the full rounds only: full_iters point-iterations x FETs device evaluations
        achieved = (tot["full_iters_1"] * n_fets * flops_per_eval / ev / 1e12) if (flops_per_eval and ev > 0) else None
        roofline = {"bound": "fp64", "kernel": "k_eval_bsimcmg107_*", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic,
                    "peak_source": "FP64 DFMA microbenchmark run live in this process (MEASURED_PEAKS.json has no FP64 figure; its "
                                   f"hbm_gbs = {peaks.get('hbm_gbs')} is the denominator for the HBM-bound k_lu / k_control)",
                    "how": "one extra step of the same workload on a single-lane plan with CUDA events around every launch "
                           "(kernels timed alone); the timed region runs the plan's lanes concurrently",
                    "flops_per_device_eval": flops_per_eval, "device_evals": tot["full_iters_1"] * n_fets,
                    "kernel_seconds": ev, "share_of_step": ev / solve1, "single_lane_step_seconds": solve1,
                    "k_lu_control_seconds": nw, "k_lu_control_share": nw / solve1,
                    "value_rounds": {"rounds": tot["value_rounds_1"], "of_rounds": tot["rounds_1"],
                                     "point_iterations": tot["newton_iters_1"] - tot["full_iters_1"],
                                     "k_evalv_seconds": tot["evalv_seconds_1"], "k_lu_solve_control_seconds": tot["newtonv_seconds_1"],
                                     "share_of_step": (tot["evalv_seconds_1"] + tot["newtonv_seconds_1"]) / solve1}}
        cpu = None
        if not args.no_cpu_baseline and world >= 1:
            threads = os.cpu_count() or 1
            pts = max(threads * 16, 64)   # a bounded sample: ~10 s of CPU work
            cel, cstats, cok = cpu_arm(pts, threads)
            cpu = {"value": pts / cel, "unit": "points/s", "cores": threads, "kind": "port",
                   "sample": f"{pts} of the {B} Monte-Carlo points, same tolerances and outputs, {cel:.1f}s",
                   "newton_iters_per_s": cstats["newton_iters"] / cel}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps({
            "metric": "transient sweep points/s (DFF Monte-Carlo)", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_gpu": B, "unknowns": fc.n_unknowns, "fets": n_fets, "swept_params": len(fc.param_names), "lanes_per_gpu": plan.lanes,
                       "l2": "per-round working set (cached device constants + device outputs, > 1 GB) exceeds the 126 MB L2; no explicit flush"},
            "newton_iters_per_s": newton_all / el, "newton_iters_per_step": newton_all / args.steps,
            "lu": circuit.lu_info(),
            "e2e": {"value": args.steps * B * world / el_e2e, "unit": "points/s", "h2d_bytes_per_step": int(P.nbytes + ts.nbytes),
                    "d2h_bytes_per_step": int(y_host.nbytes + status.nbytes)},
            "gpu_launches": int(launches_all),
            "clocks": sampler.summary(),
            "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": {"converged_points": ok_points, "of": B, "max_abs_q_error_vs_known_pattern_V": max(q_known)},
        }), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    plan.close()


if __name__ == "__main__":
    main()
