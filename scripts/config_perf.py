"""Throughput of the other BASELINE.json configurations at their full sizes on one GPU (the bench line is config 3):
config 4 (BSIM-CMG I-V sweep, 1 048 576 DC bias points), config 2 (inverter VDD x NFIN x L product sweep, 65 536 points,
fixed-step trapezoidal, 8 000 steps), config 5 stand-in (corner x temperature x mismatch, 131 072 adaptive transients).
One warm pass, one timed pass each; device time of the solve from the engine's CUDA events.

    python scripts/config_perf.py > gpurun_out/config_perf.log
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cedarsim.jl_b200 import circuits, engine  # noqa: E402
from cedarsim.jl_b200.flat import params_matrix  # noqa: E402


def report(name, B, stats, wall):
    s = stats["solve_seconds"]
    print(json.dumps({"config": name, "points": B, "solve_seconds": s, "wall_seconds": wall, "points_per_s": B / s,
                      "newton_iters_per_s": stats["newton_iters"] / s, "newton_iters_per_point": stats["newton_iters"] / B,
                      "rounds": stats["rounds"], "kernel_launches": stats["kernel_launches"]}), flush=True)


def main():
    # config 4
    fc, ms = circuits.fet_iv()
    n = 1024
    vg, vd = np.meshgrid(np.linspace(0, 0.9, n), np.linspace(0, 0.9, n), indexing="ij")
    P = params_matrix([vg.ravel(order="F"), vd.ravel(order="F")])
    plan = engine.Circuit(fc, ms).plan(P.shape[1])
    plan.set_params(P)
    plan.dc()
    t = time.perf_counter()
    _, _, st, stats = plan.dc()
    report("4: BSIM-CMG I-V sweep, DC operating points", P.shape[1], stats, time.perf_counter() - t)
    assert st.max() == 0
    plan.close()
    # config 2
    fc, ms = circuits.inverter()
    vdd, nfin, ln = np.meshgrid(np.linspace(0.8 * 0.7, 1.2 * 0.7, 16), np.linspace(3.0, 12.0, 64), np.linspace(21e-9, 63e-9, 64), indexing="ij")
    B = vdd.size
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = vdd.ravel(order="F")
    P[fc.param_names.index("xneg.nfin")] = nfin.ravel(order="F")
    P[fc.param_names.index("xneg.l")] = ln.ravel(order="F")
    ts = np.linspace(0, 4e-7, 401)
    opts = engine.default_options(fixed_step=1, dt=50e-12)
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    t = time.perf_counter()
    _, st, stats = plan.tran(0.0, 4e-7, ts, opts)
    report("2: inverter product sweep, fixed-step trapezoidal, 8 000 steps", B, stats, time.perf_counter() - t)
    assert st.max() == 0
    plan.close()
    # config 5 stand-in
    from cedarsim.jl_b200.sweeps import CircuitSweep, tran_
    cs = CircuitSweep(circuits.CONFIG5_DECK, circuits.config5_sweep(), outputs=["q", "d", "vvdd.i"])
    ts = np.linspace(0, 2.5e-9, 251)
    tran_(cs, (0.0, 2.5e-9), saveat=ts, reltol=1e-3)
    t = time.perf_counter()
    sols = tran_(cs, (0.0, 2.5e-9), saveat=ts, reltol=1e-3)
    report("5 (stand-in): corner x temperature x mismatch transient, adaptive", len(cs), sols.stats, time.perf_counter() - t)
    assert sols.status.max() == 0


if __name__ == "__main__":
    main()
