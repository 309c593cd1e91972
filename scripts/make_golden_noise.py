"""Extracts the ngspice output-noise tables that the reference's own tests assert against
(test/ac.jl:67-149 Butterworth filter, test/ac.jl:161-237 BSIM-CMG inverter) into tests/golden/.
Run where /root/reference exists; the fixtures are committed because the GPU box has no reference tree."""
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "test/ac.jl")).read()
tables = re.findall(r"ngspice = \[\n(.*?)\n\]", src, re.S)
assert len(tables) == 2
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(out, exist_ok=True)
for name, cite, tab in (("ngspice_noise_butterworth.txt", "test/ac.jl:85-147", tables[0]),
                        ("ngspice_noise_bsimcmg_inverter.txt", "test/ac.jl:173-235", tables[1])):
    rows = [l.split() for l in tab.strip().splitlines()]
    with open(os.path.join(out, name), "w") as f:
        f.write(f"# frequency_Hz  onoise_V_per_sqrtHz   (ngspice table asserted by the reference, {cite}, rtol 1e-6)\n")
        for r in rows:
            f.write(f"{r[0]} {r[1]}\n")
    print(name, len(rows))
