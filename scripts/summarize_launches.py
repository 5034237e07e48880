"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the captured window)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
data = rows[rows.index(hdr) + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else v * 1000 if r[mu] == "ms" else v
    name = re.sub(r"\(.*", "", r[kn])[:70]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"# total {tot / 1000:.2f} ms over {len(data)} launches")
print("kernel,launches,total_us,share")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k},{n},{t:.1f},{t / tot:.3f}")
