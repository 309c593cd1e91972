"""Per-kernel launch statistics of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
    agg[row["Kernel Name"].split("(")[0]].append((us, row.get("Grid Size", "")))
tot = sum(sum(u for u, _ in v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(u for u, _ in kv[1])):
    us = [u for u, _ in v]
    print(f"{k:42s} n={len(v):4d} mean={sum(us) / len(us):8.1f} us  min={min(us):7.1f} max={max(us):7.1f}  share={100 * sum(us) / tot:5.1f} %  grid={v[0][1]}")
