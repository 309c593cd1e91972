"""Stall samples and executed instructions per SASS opcode from `ncu --page source --csv --print-source sass`; sections of
out-of-line functions are reported too (by the address gap)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(r for r in rows if "# Samples" in r); ix = {x: i for i, x in enumerate(h)}
tot = 0; by = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter(); data = []
for r in rows:
    if len(r) < len(h) or r is h: continue
    try: s = float(r[ix["# Samples"]]); n = float(r[ix["Instructions Executed"]])
    except ValueError: continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0] if parts else "?"
    by[op] += s; cnt[op] += 1; ex[op] += n; tot += s
    data.append((s, src, n))
dyn = sum(ex.values())
print("total samples", tot, "static instr", len(data), "dyn warp instr", dyn)
for op, s in by.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} samples {100*s/tot:5.1f}%  static {cnt[op]:5d}  dyn {100*ex[op]/dyn:5.1f}%")
print("--- top instructions")
for s, src, n in sorted(data, reverse=True)[:25]: print(f"{100*s/tot:5.2f}% {src[:100]}")
