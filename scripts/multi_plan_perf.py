"""SURVEY.md 8(e) "measure both": the DFF Monte-Carlo job (16 384 points in total) on the GPUs of ONE process through
cb_plan_create_multi -- one host thread per lane and device inside the library, every device copies its slice of the
waveforms straight into the caller's pinned host array, no NCCL -- next to the torchrun + NCCL-gather figure of bench.py.

    python scripts/multi_plan_perf.py 2        # number of GPUs
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cedarsim.jl_b200 import circuits, engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    fc, ms = circuits.dff()
    circuit = engine.Circuit(fc, ms)
    B = bench.TOTAL_POINTS
    P = circuits.dff_mc_params(fc, B)
    ts = np.linspace(bench.T0, bench.T1, bench.NSAVE)
    opts = engine.default_options(**bench.OPTS, **bench.ENGINE_OPTS)
    plan = circuit.plan(B, devices=list(range(n)))
    plan.set_x0(bench.nodeset(fc))
    O = len(fc.outputs)
    y = torch.empty((O, bench.NSAVE, B), dtype=torch.float64, pin_memory=True).numpy()
    Ph = torch.from_numpy(P).pin_memory().numpy()
    plan.set_params(Ph)
    plan.tran(bench.T0, bench.T1, ts, opts, out=y)
    steps = 3
    t = time.perf_counter()
    for _ in range(steps):
        plan.set_params(Ph)
        y, st, stats = plan.tran(bench.T0, bench.T1, ts, opts, out=y)
    el = (time.perf_counter() - t) / steps
    print(json.dumps({"gpus": n, "plan_devices": plan.n_devices, "lanes": plan.lanes, "points": B, "seconds_per_step": el,
                      "e2e_points_per_s": B / el, "converged": int((st == 0).sum())}), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
