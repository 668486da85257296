"""ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_transpose.py [regex] > summary.csv
One line per metric, one column per captured launch (optionally only metrics matching regex)."""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
w = csv.writer(sys.stdout, quoting=csv.QUOTE_MINIMAL)
w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
for j, name in enumerate(hdr):
    if name in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device",
                "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"):
        continue
    if pat and not pat.search(name) and name != "Kernel Name":
        continue
    w.writerow([name, units[j]] + [r[j] if j < len(r) else "" for r in data])
