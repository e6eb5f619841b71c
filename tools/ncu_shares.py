"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_shares.py launches.csv ["header line"]"""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    name = r[ik].split("(")[0].replace("void dbx::", "").replace("dbx::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("total %.1f us over %d launches (cold-cache, serialised: compare SHARES)" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005:
        continue
    print("%-52s %4d launches %10.1f us %5.1f%%" % (k[:52], v[0], v[1], 100 * v[1] / tot))
