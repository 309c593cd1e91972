"""NVRTC-compile the DFF engine for a list of experiment variants (cubins land in the cache that travels to the GPU box)."""
import os, sys, subprocess
variants = [v.split() for v in open(sys.argv[1]).read().strip().splitlines() if v.strip() and not v.startswith("#")]
procs = []
for v in variants:
    env = dict(os.environ, **dict(kv.split("=", 1) for kv in v))
    procs.append(subprocess.Popen([sys.executable, "-c", "import sys; sys.path.insert(0,'.'); from cedarsim.jl_b200 import circuits, engine; fc, ms = circuits.dff(); engine.Circuit(fc, ms)"], env=env))
for p in procs: p.wait()
