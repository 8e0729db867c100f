"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("ngicp::<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':44s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / 1e3 / v[0]:9.2f} {100 * v[1] / tot:6.1f}%")
print(f"{'TOTAL':44s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.1f}")
