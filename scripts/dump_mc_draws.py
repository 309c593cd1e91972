"""Write the Monte-Carlo draws of the bench workload (circuits.dff_mc_params, seed 20240607) as a CSV whose header holds
the swept names: the input of bench/cedarsim_cpu_sweep.jl, so that the reference integrates the same instances.

    python scripts/dump_mc_draws.py [points] [out.csv]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cedarsim.jl_b200 import circuits  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = sys.argv[2] if len(sys.argv) > 2 else "mc_draws.csv"
fc, _ = circuits.dff()
P = circuits.dff_mc_params(fc, points)
with open(out, "w") as fh:
    fh.write(",".join(fc.param_names) + "\n")
    for b in range(points):
        fh.write(",".join(repr(float(v)) for v in P[:, b]) + "\n")
print(out, P.shape)
