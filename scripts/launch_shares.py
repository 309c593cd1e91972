"""Share of the serialised time per kernel family in an ncu launch list (gpu__time_duration.sum CSV).
usage: python scripts/launch_shares.py profiles/launches_r1ak.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = r[kn]
    fam = ("k_evalv_*" if name.startswith("k_evalv_") else "k_eval_*" if name.startswith("k_eval_") else re.sub(r"[<(].*", "", name))
    tot[fam] += float(r[mv].replace(",", ""))
    cnt[fam] += 1
s = sum(tot.values())
for fam in sorted(tot, key=tot.get, reverse=True):
    print(f"{fam:24s} launches {cnt[fam]:5d}  mean {tot[fam] / cnt[fam] / 1e3:8.1f} us  share {100 * tot[fam] / s:5.1f} %")
