"""Summarise `ncu --page raw --csv` dumps: per launch duration, DRAM bytes, tensor-pipe activity, L2->SM traffic.
usage: python tools/ncu_summary.py raw.csv [more.csv ...]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM"),
        ("launch__registers_per_thread", "regs")]


def unit_scale(u, target):
    table = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3,
             "Gbyte/s": 1e-3, "Tbyte/s": 1.0, "Mbyte/s": 1e-6}
    return table.get(u, 1.0)


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    print("== %s (%d launches)" % (path, len(body)))
    print("%-34s %8s %8s %8s %9s %9s %6s %8s %5s" % (("kernel",) + tuple(c[1] for c in COLS)))
    tot = [0.0, 0.0, 0.0]
    for r in body:
        vals = []
        for name, _ in COLS:
            i = idx.get(name)
            if i is None:
                vals.append(float("nan"))
                continue
            vals.append(float(r[i].replace(",", "")) * unit_scale(units[i], name))
        tot[0] += vals[0]; tot[1] += vals[1]; tot[2] += vals[2]
        name = r[kn].split("(")[0].replace("void dbx::", "").replace("(bool)", "")[:34]
        print("%-34s %8.1f %8.1f %8.1f %9.1f %9.1f %6.1f %8.2f %5.0f" % ((name,) + tuple(vals)))
    print("total: %.1f us, DRAM read %.1f MB, write %.1f MB -> per launch %.1f MB" % (
        tot[0], tot[1], tot[2], (tot[1] + tot[2]) / max(len(body), 1)))
