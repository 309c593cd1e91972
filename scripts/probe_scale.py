"""Throughput of the bench workload at several batch sizes / lane counts on one GPU (what a GPU sees under strong
scaling of BASELINE config 3: 16 384 points total over 1/2/4/8 GPUs = 16 384 / 8 192 / 4 096 / 2 048 points each).

    python scripts/probe_scale.py [B:lanes ...]      # default: 2048:1 2048:2 4096:2 8192:4 16384:4
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cedarsim.jl_b200 import circuits, engine  # noqa: E402


def main():
    cases = sys.argv[1:] or ["2048:1", "2048:2", "4096:2", "8192:4", "16384:4"]
    fixed = os.environ.get("PROBE_FIXED")   # "dt" in seconds -> fixed-step mode
    fc, ms = circuits.dff()
    circuit = engine.Circuit(fc, ms)
    ts = np.linspace(bench.T0, bench.T1, bench.NSAVE)
    kw = dict(bench.OPTS, **bench.ENGINE_OPTS)
    if fixed:
        kw.update(fixed_step=1, dt=float(fixed))
    opts = engine.default_options(**kw)
    for case in cases:
        B, lanes = (int(x) for x in case.split(":"))
        plan = circuit.plan(B, device=0, lanes=lanes)
        plan.set_x0(bench.nodeset(fc))
        plan.set_params(circuits.dff_mc_params(fc, B))
        plan.tran_device(bench.T0, bench.T1, ts, opts)
        t = time.perf_counter()
        reps = 2
        for _ in range(reps):
            _, _, st = plan.tran_device(bench.T0, bench.T1, ts, opts)
        el = (time.perf_counter() - t) / reps
        print(json.dumps({"B": B, "lanes": plan.lanes, "fixed": fixed, "sec": el, "points_per_s": B / el, "rounds": st["rounds"],
                          "value_rounds": st["value_rounds"], "us_per_round": 1e6 * el / max(1, st["rounds"]) * plan.lanes,
                          "iters_per_point": st["newton_iters"] / B, "full_iters_per_point": st["full_iters"] / B,
                          "launches": st["kernel_launches"],
                          "timing_s": {k: round(st[k], 4) for k in ("eval_seconds", "evalv_seconds", "newton_seconds", "newtonv_seconds", "solve_seconds")}}), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
