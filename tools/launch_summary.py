"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [out_summary.csv] [--last N]   (--last N: only the last N launches)"""
import collections
import csv
import sys

last = 0
if "--last" in sys.argv:
    i = sys.argv.index("--last")
    last = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names = rows[hdr]
kn, mv, mu = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
agg = collections.OrderedDict()
body = [r for r in rows[hdr + 1:] if len(r) > mv]
for r in (body[-last:] if last else body):
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000.0 if r[mu] == "ns" else (v * 1000.0 if r[mu] == "ms" else v)
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out = [["kernel", "launches", "total_us", "avg_us", "share"]]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.2f}", f"{a[1] / tot:.4f}"])
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w", newline="")).writerows(out)
for r in out:
    print(",".join(str(x) for x in r))
print("total_us", round(tot, 1), "launches", sum(a[0] for a in agg.values()))
