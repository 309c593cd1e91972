"""Offline resource / SASS statistics of the generated device kernels of the DFF circuit (no GPU needed):
registers, spill bytes, static SASS instruction count and mix per kernel.

    [CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=2] [CB_GEN_DIR=_gen_x] python scripts/eval_stats.py [kernel-substring]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cedarsim.jl_b200 import circuits, engine  # noqa: E402

want = sys.argv[1] if len(sys.argv) > 1 else "k_eval"
d = tempfile.mkdtemp()
fc, ms = circuits.dff()
engine.Circuit(fc, ms, cache_dir=d)
cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    if want in m.group(1):
        print(m.group(1), m.group(2))
sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
cur, mix = None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        mix[cur][m.group(1).split(".")[0]] += 1
for k, c in mix.items():
    if want in k:
        tot = sum(c.values())
        fp64 = sum(v for op, v in c.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
        print(f"{k}: {tot} SASS instructions, FP64-class {fp64} ({100 * fp64 / tot:.0f} %); top: " +
              ", ".join(f"{op} {v}" for op, v in c.most_common(14)))
