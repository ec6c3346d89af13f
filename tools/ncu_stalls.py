"""Top stall sites of an `ncu --page source --csv` dump (SASS view): usage: python tools/ncu_stalls.py source.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ix = {n: i for i, n in enumerate(names)}
stall_cols = [n for n in names if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hdr + 1:]:
    if len(r) < len(names):
        continue
    try:
        ns = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((ns, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
agg = {c: 0 for c in stall_cols}
for ns, r in data:
    for c in stall_cols:
        try:
            agg[c] += int(r[ix[c]] or 0)
        except ValueError:
            pass
print("by reason:", ", ".join(f"{c[6:]}={v}" for c, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
for ns, r in sorted(data, key=lambda d: -d[0])[:top]:
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{ns:7d} {100.0 * ns / max(tot, 1):5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:70]:70s} " +
          " ".join(f"{n}:{v}" for v, n in reasons if v))
