"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import csv, sys, re
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    unit, val = r[13], float(r[14].replace(",", ""))
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
allns = sum(v[1] for v in tot.values())
print("%-60s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %6.2f%%" % (k[:60], v[0], v[1] / 1e3, 100 * v[1] / allns))
print("%-60s %8d %12.1f" % ("TOTAL", sum(v[0] for v in tot.values()), allns / 1e3))
