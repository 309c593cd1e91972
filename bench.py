#!/usr/bin/env python
"""bench.py -- headline benchmark: transient sweep points/s on the DFF Monte-Carlo workload.

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm on the box's host cores

One "step" = one full pass of the hot path over one batch: DC operating point + adaptive
transient (0 .. 600 ns, LTE-controlled trapezoidal) of the 16 384 Monte-Carlo instances of the
30-FET D flip-flop (BASELINE.json configs[2] as written: 16 384 instances TOTAL at 1/2/4/8 GPUs, i.e.
STRONG scaling; GF180/BSIM4 are not in the reference tree, so the same topology runs on BSIM-CMG 107 +
ASAP7 cards -- SURVEY.md fact 5).

`value` is timed with inputs resident in HBM and results left in HBM; `e2e` goes through the
C-ABI calls with pinned host buffers (H2D of the parameter matrix and D2H of all waveforms inside
the timed region).  Under torchrun rank r owns the contiguous block [r B/N, (r+1) B/N) of the same 16 384 draws (no
data-path collective) and rank 0 gathers the waveforms over NCCL.  Extra legs on the same JSON line: `fixed_step`
(config 3's second timing mode, dt = 25 ps), `weak` (N > 1: 16 384 points per GPU), `compile_seconds` (N = 1: cold
NVRTC + symbolic analysis; the Verilog-A code generation time recorded at build()).
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
TOTAL_POINTS = 16384          # BASELINE config 3: instances in the whole job, whatever the number of GPUs
POINTS_PER_GPU = TOTAL_POINTS   # weak-scaling leg and the single-GPU case
FIXED_DT = 25e-12             # config 3's fixed-step comparison mode
# SURVEY.md 8(d) config 3, adaptive mode: reltol 1e-4, abstol 1e-6 V / 1e-12 A.  Newton tolerance = 0.1 x the LTE
# tolerance with the contraction-rate acceptance test (Sundials IDA, the reference's solver, uses 0.33 x and the same
# test); the CPU arm runs the same options.  value_rounds is the engine's chord-iteration schedule.
OPTS = dict(reltol=1e-4, vabstol=1e-6, iabstol=1e-12, nr_reltol=1e-5, nr_vabstol=1e-7, nr_iabstol=1e-13, nr_rate_test=1)
ENGINE_OPTS_BASE = dict(value_rounds=2)                       # chord-iteration cycle: restated by the oracle as well
# engine-side schedule of the same iterations: lock-step rounds (mixed_rounds=1 takes 27 % fewer rounds but measured slower
# on this workload, profiles/probe_r2*.log)
ENGINE_OPTS = dict(ENGINE_OPTS_BASE, mixed_rounds=0)
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


def cpu_arm(points, threads, fast=True):
    """The CPU restatement on `points` sweep points with `threads` host threads.
    fast=True : `cpu_fast` -- same equations and step control, static-pivot sparse LU (the engine's symbolic analysis),
                -O3 -march=native for the solver and the generated device models: the honest CPU baseline.
    fast=False: the CHECKER as it is used in the parity tests (dense partial-pivoting LU, -O2, no FP contraction)."""
    from cedarsim.jl_b200 import circuits
    from oracle import orc

    def run():
        fc, _ = circuits.dff(host=("fast:" + orc.cpu_tag()) if fast else True)
        P = circuits.dff_mc_params(fc, TOTAL_POINTS)[:, :points]
        P = np.ascontiguousarray(P)
        orc.set_x0(nodeset(fc))
        ts = np.linspace(T0, T1, NSAVE)
        t = time.perf_counter()
        # the same algorithm as the engine: chord iterations (value_rounds) with the derivative-free device functions
        y, st, stats = orc.tran(fc, T0, T1, ts, params=P, opts=orc.default_options(**OPTS, **ENGINE_OPTS_BASE), nthreads=threads)
        el = time.perf_counter() - t
        orc.set_x0(None)
        return el, stats, int((st == 0).sum())
    if fast:
        with orc.fast_arm():
            return run()
    return run()


def cpu_sample_points(threads, big=False):
    """bounded samples of the 16 384 draws: ~10 s of CPU work per step for the reference arm (many steps), 2 048 points
    (~1 min on 16 cores) for the cpu_baseline figure of the engine's own line"""
    return 2048 if big else max(threads * 16, 256)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation cannot run here (Julia, un-vendored packages), so this is
    the CPU restatement at its fastest (`cpu_fast`) on all host threads, on a bounded sample of the same draws."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    points = cpu_sample_points(threads)
    for _ in range(min(args.warmup, 1)):   # builds the -march=native libraries on first use
        cpu_arm(min(points, threads * 4), threads)
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_step": points, "total_points": TOTAL_POINTS},
        "newton_iters_per_s": iters / el,
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "arm": "cpu_fast: chord iterations with value-only device functions, sparse static-pivot LU, -O3 -march=native, OpenMP over points",
                         "sample": f"the first {points} of the {TOTAL_POINTS} Monte-Carlo points per step, same tolerances and outputs"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def codegen_seconds(ms):
    """Verilog-A -> CUDA C generation time of the models, recorded when build() generated them (the model sources are
    not on the GPU box)."""
    try:
        return float(sum(getattr(cm, "codegen_seconds", 0.0) for cm in ms))
    except TypeError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--points", type=int, default=0, help="sweep points per GPU (default: 16384 / number of GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the fixed-step, weak-scaling and cold-compile legs")
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
    B = args.points or TOTAL_POINTS // world            # strong scaling: the job is 16 384 points whatever N is
    fc, ms = circuits.dff()
    circuit = engine.Circuit(fc, ms)
    plan = circuit.plan(B, device=local)
    P_all = circuits.dff_mc_params(fc, max(TOTAL_POINTS, B * world))
    P = np.ascontiguousarray(P_all[:, rank * B:(rank + 1) * B])
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

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def resident_leg(pl, Bl, o, warm, steps, sample_clocks=False):
        """`steps` timed passes of plan `pl` (Bl points per GPU), inputs and results in HBM, waveforms gathered on rank 0."""
        gbuf = [torch.empty((O, NSAVE, Bl), dtype=torch.float64, device="cuda") for _ in range(world)] if (world > 1 and rank == 0) else None

        def step():
            dy, ds, st = pl.tran_device(T0, T1, ts, o)
            if world > 1:   # final waveforms to rank 0 over NVLink (NCCL); no collective on the solve path
                dist.gather(torch.as_tensor(DevArray(dy, (O, NSAVE, Bl)), device="cuda"), gbuf, dst=0)
            return st
        for _ in range(warm):
            step()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        tot = {}
        for _ in range(steps):
            st = step()
            for k, v in st.items():
                tot[k] = tot.get(k, 0) + v
        barrier()
        el = allmax(time.perf_counter() - t0)
        if sampler:
            sampler.stop_flag.set()
        return el, tot, sampler

    el, tot, sampler = resident_leg(plan, B, opts, args.warmup, args.steps, sample_clocks=True)
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
    el_e2e = allmax(time.perf_counter() - t0)
    ok_points = int((status == 0).sum())
    q_known = [float(np.abs(y_host[0, int(round(tt / T1 * (NSAVE - 1)))] - want).max())
               for tt, want in ((1.5e-7, 0.0), (2.5e-7, 0.0), (4.5e-7, 0.7), (5.5e-7, 0.7))]

    # ---- config 3's second timing mode: fixed step, dt = 25 ps (24 000 steps per point), one timed pass
    fixed = None
    if not args.no_extra_legs:
        fopts = engine.default_options(**OPTS, **ENGINE_OPTS, fixed_step=1, dt=FIXED_DT)
        elf, totf, _ = resident_leg(plan, B, fopts, 0, 1)
        _, ds_f, _ = 0, plan.last_status_ptr, 0
        st_f = torch.as_tensor(DevArray(ds_f, (B,), "<i4"), device="cuda").cpu().numpy() if ds_f else None
        fixed = {"value": B * world / elf, "unit": "points/s", "dt": FIXED_DT, "ms_per_step": 1e3 * elf, "steps": 1,
                 "converged_points_this_rank": int((st_f == 0).sum()) if st_f is not None else None,
                 "rounds": totf["rounds"],
                 "newton_iters_per_point": totf["newton_iters"] / B, "accepted_steps_per_point": totf["steps_accepted"] / B}
    # ---- weak scaling next to the strong-scaling headline: 16 384 points on EVERY GPU
    weak = None
    if world > 1 and not args.no_extra_legs and not args.points:
        planw = circuit.plan(POINTS_PER_GPU, device=local)
        planw.set_x0(nodeset(fc))
        planw.set_params(np.ascontiguousarray(circuits.dff_mc_params(fc, POINTS_PER_GPU, seed=20240607 + rank)))
        elw, _, _ = resident_leg(planw, POINTS_PER_GPU, opts, 1, 2)
        planw.close()
        weak = {"value": 2 * POINTS_PER_GPU * world / elw, "unit": "points/s", "points_per_gpu": POINTS_PER_GPU, "steps": 2,
                "ms_per_step": 1e3 * elw / 2}
    # ---- compile latency (north star: reported separately): cold NVRTC + symbolic analysis, no cubin cache
    compile_s = None
    if rank == 0 and world == 1 and not args.no_extra_legs:
        compile_s = {"nvrtc_and_symbolic_cold": engine.Circuit(fc, ms, cache_dir=None).compile_seconds,
                     "with_cubin_cache": circuit.compile_seconds, "verilog_a_codegen_at_build": codegen_seconds(ms)}

    if rank == 0:
        # ---- roofline of the dominant kernel (device evaluation): both roofs, the binding one is the slower floor
        flops_per_eval = float(np.mean([sum(cm.exec_ops) for cm in ms])) if all(getattr(cm, "exec_ops", None) for cm in ms) else None
        n_fets = len(fc.va_insts)
        fp64_peak = engine.measure_fp64_peak(local)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs") or 6557.8)   # fallback: the figure of B200_PROFILING.md's recipe on this pool
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("k_eval_dram_bytes_per_launch")
        except OSError:
            pass
        ev, nw = tot["eval_seconds_1"], tot["newton_seconds_1"]
        solve1 = max(tot["solve_seconds_1"], 1e-30)
        # k_eval_* runs in the full iterations only: full_iters point-iterations x FETs device evaluations.
        # Algorithmic bytes per evaluation: the per-instance cache row (ncache slots), the terminal voltages, the outputs
        # (I, Q per terminal; J and C per Jacobian entry).
        evals = tot["full_iters_1"] * n_fets
        bytes_per_eval = float(np.mean([8 * (cm.ncache + len(cm.terminals) + 2 * len(cm.terminals) + 2 * len(cm.jrow)) for cm in ms]))
        tf = (evals * flops_per_eval / ev / 1e12) if (flops_per_eval and ev > 0) else None
        gbs = evals * bytes_per_eval / ev / 1e9 if ev > 0 else None
        frac_fp64 = tf / fp64_peak if tf else None
        frac_hbm = gbs / hbm_peak if gbs else None
        bound = "hbm" if (frac_hbm or 0) > (frac_fp64 or 0) else "fp64"
        roofline = {"bound": bound, "kernel": "k_eval_bsimcmg107_*",
                    "achieved": gbs if bound == "hbm" else tf, "peak": hbm_peak if bound == "hbm" else fp64_peak,
                    "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": frac_hbm if bound == "hbm" else frac_fp64,
                    "traffic": traffic,
                    "fp64": {"achieved_tflops": tf, "peak_tflops": fp64_peak, "frac": frac_fp64, "flops_per_device_eval": flops_per_eval},
                    "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": frac_hbm, "bytes_per_device_eval": bytes_per_eval},
                    "peak_source": "fp64: DFMA microbenchmark run live in this process (MEASURED_PEAKS.json has no FP64 figure); "
                                   f"hbm: MEASURED_PEAKS.json hbm_gbs = {peaks.get('hbm_gbs')}; the binding roof is the one with the larger fraction",
                    "how": "one extra step of the same workload on a single-lane plan with CUDA events around every launch "
                           "(kernels timed alone); the timed region runs the plan's lanes concurrently",
                    "device_evals": evals, "kernel_seconds": ev, "share_of_step": ev / solve1, "single_lane_step_seconds": solve1,
                    "k_lu_control_seconds": nw, "k_lu_control_share": nw / solve1,
                    "value_rounds": {"rounds": tot["value_rounds_1"], "of_rounds": tot["rounds_1"],
                                     "point_iterations": tot["newton_iters_1"] - tot["full_iters_1"],
                                     "k_evalv_seconds": tot["evalv_seconds_1"], "k_lu_solve_control_seconds": tot["newtonv_seconds_1"],
                                     "share_of_step": (tot["evalv_seconds_1"] + tot["newtonv_seconds_1"]) / solve1}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            pts = cpu_sample_points(threads, big=True)
            cpu_arm(threads * 2, threads)        # builds the -march=native libraries on first use, untimed
            cel, cstats, cok = cpu_arm(pts, threads)
            dpts = max(threads * 4, 32)
            del_, dstats, _ = cpu_arm(dpts, threads, fast=False)
            cpu = {"value": pts / cel, "unit": "points/s", "cores": threads, "kind": "port",
                   "arm": "cpu_fast: same equations, step control and chord iterations (value-only device functions), static-pivot sparse LU "
                          "(engine's symbolic analysis), -O3 -march=native, OpenMP over points",
                   "sample": f"the first {pts} of the {TOTAL_POINTS} Monte-Carlo points, same tolerances and outputs, {cel:.1f}s",
                   "newton_iters_per_s": cstats["newton_iters"] / cel,
                   "checker_dense": {"value": dpts / del_, "unit": "points/s", "sample": f"{dpts} points, {del_:.1f}s",
                                     "note": "the parity checker as the tests use it: dense partial-pivoting LU, -O2, no FP contraction"}}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps({
            "metric": "transient sweep points/s (DFF Monte-Carlo)", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "strong" if not args.points else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "total_points": B * world, "points_per_gpu": B, "unknowns": fc.n_unknowns, "fets": n_fets,
                       "swept_params": len(fc.param_names), "lanes_per_gpu": plan.lanes, "engine_options": ENGINE_OPTS,
                       "l2": "per-round working set (cached device constants + device outputs, > 1 GB at 16 384 points) exceeds the 126 MB L2; no explicit flush"},
            "newton_iters_per_s": newton_all / el, "newton_iters_per_step": newton_all / args.steps,
            "lu": circuit.lu_info(),
            "e2e": {"value": args.steps * B * world / el_e2e, "unit": "points/s", "h2d_bytes_per_step": int(P.nbytes + ts.nbytes),
                    "d2h_bytes_per_step": int(y_host.nbytes + status.nbytes)},
            "gpu_launches": int(launches_all),
            "clocks": sampler.summary(),
            "roofline": roofline, "cpu_baseline": cpu,
            "fixed_step": fixed, "weak": weak, "compile_seconds": compile_s,
            "parity_check": {"converged_points": ok_points, "of": B, "max_abs_q_error_vs_known_pattern_V": max(q_known)},
        }), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    plan.close()


if __name__ == "__main__":
    main()
