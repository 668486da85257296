"""Per-kernel totals of an ncu launch list: python scripts/launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

tot = OrderedDict()
with open(sys.argv[1]) as fh:
    rows = [r for r in csv.reader(fh) if len(r) > 14 and r[0].isdigit()]
for r in rows:
    name = re.sub(r"^void ", "", r[4]).split("(")[0]
    val = float(r[14].replace(",", ""))
    unit = r[13]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + us)
total = sum(t for _, t in tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (name, n, t, t / n, 100 * t / total))
