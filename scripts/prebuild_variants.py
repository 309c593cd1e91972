import os, sys, subprocess
variants = [dict(), dict(CB_NVRTC_DEFS="-DVA_EVAL_MINBLOCKS=10"), dict(CB_NVRTC_DEFS="-DVA_EVAL_MINBLOCKS=12"), dict(CB_NVRTC_DEFS="-DVA_EVAL_MINBLOCKS=6")]
procs = []
for v in variants:
    env = dict(os.environ, **v)
    procs.append(subprocess.Popen([sys.executable, "-c", "import sys; sys.path.insert(0,'.'); from cedarsim.jl_b200 import circuits, engine; fc, ms = circuits.dff(); engine.Circuit(fc, ms)"], env=env))
for p in procs: p.wait()
