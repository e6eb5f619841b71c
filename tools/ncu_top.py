"""Top stall-sampled SASS instructions of an `ncu --page source --csv` dump (with a little context)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
iexec = hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > isamp]
tot = sum(int(r[isamp] or 0) for r in body)
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
print("total samples", tot, "instructions", len(body))
for i in sorted(order):
    r = body[i]
    print("%5d %6.2f%% exec=%8s  %s" % (i, 100.0 * int(r[isamp]) / max(tot, 1), r[iexec], r[isrc].strip()[:110]))
