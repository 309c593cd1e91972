"""Throughput of the small-signal path: output noise (cb_noise) and AC response (cb_ac) of the 30-FET DFF for B sweep
points x F frequencies; prints device seconds of the linearisation + complex LU and (point, frequency) systems / s."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import bench
from cedarsim.jl_b200 import circuits, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = int(sys.argv[2]) if len(sys.argv) > 2 else 61
fc, ms = circuits.dff()
fc.set_outputs(["q"])
P = circuits.dff_mc_params(fc, B)
c = engine.Circuit(fc, ms)
p = c.plan(B); p.set_params(P); p.set_x0(bench.nodeset(fc))
f = 10.0 ** np.linspace(3, 12, F)
for name, fn in (("noise", p.noise), ("ac", p.ac)):
    for rep in range(2):
        t = time.time(); out, st, stats = fn(f); el = time.time() - t
        print(f"{name} rep{rep} B={B} F={F} lanes={p.lanes} wall {el:.3f}s small-signal device {stats['newton_seconds']:.4f}s "
              f"systems/s {B * F / max(stats['newton_seconds'], 1e-9):.3e} ok={int((st == 0).sum())}", flush=True)
