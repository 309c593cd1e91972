"""Per-source-line stall samples / executed instructions from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(r for r in rows if "# Samples" in r); ix = {x: i for i, x in enumerate(h)}
agg = {}
for r in rows:
    if len(r) < len(h) or not r[0] or not r[0].isdigit(): continue
    try: k = int(r[0]); s, n = float(r[ix["# Samples"]]), float(r[ix["Instructions Executed"]])
    except ValueError: continue
    a = agg.setdefault(k, [0.0, 0.0, r[1][:130]]); a[0] += s; a[1] += n
tot = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("samples", tot, "warp instr", ti)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
for k in sorted(agg):
    a = agg[k]
    if a[0] / tot > thr or a[1] / ti > thr: print(f"{100*a[0]/tot:5.1f}%s {100*a[1]/ti:5.1f}%i  {k:>5} {a[2]}")
