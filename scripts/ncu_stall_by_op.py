"""Stall reasons per SASS opcode from `ncu --page source --csv --print-source sass` (which opcode waits for what)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(r for r in rows if "# Samples" in r); ix = {x: i for i, x in enumerate(h)}
reasons = [x for x in h if x.startswith("stall_") and "(Not Issued)" not in x]
agg = collections.defaultdict(lambda: collections.Counter()); tot = collections.Counter()
for r in rows:
    if len(r) < len(h) or r is h: continue
    try: float(r[ix["# Samples"]])
    except ValueError: continue
    parts = r[ix["Source"]].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0] if parts else "?"
    for k in reasons:
        v = float(r[ix[k]] or 0); agg[op][k] += v; tot[k] += v
T = sum(tot.values())
print("reasons overall:", {k: f"{100*v/T:.1f}%" for k, v in tot.most_common(9)})
for op, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    s = sum(c.values())
    print(f"{op:8s} {100*s/T:5.1f}%  " + "  ".join(f"{k[6:]}={100*v/s:.0f}%" for k, v in c.most_common(4)))
