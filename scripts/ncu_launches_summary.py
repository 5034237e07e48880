"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total / mean time."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # ignore the first `skip` launches (warm-up)
agg = OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = r[ki].replace("<unnamed>::", "").replace("void ", "")
    name = name.split("(")[0][:70]
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {c:8d} {t / 1e3:10.1f} {t / 1e3 / c:9.2f} {100 * t / tot:5.1f}%")
print(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):8d} {tot / 1e3:10.1f}")
