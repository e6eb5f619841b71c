"""Per-kernel summary of an `ncu --page source --csv` + `--page raw --csv` pair: executed warp instructions, opcode
histogram (weighted by executions), issue utilisation, occupancy, waves.  usage: python tools/ncu_ops.py <tag>
(reads gpurun_out/<tag>_src.csv and gpurun_out/<tag>_raw.csv)"""
import csv
import sys

tag = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = list(csv.reader(open("gpurun_out/%s_raw.csv" % tag)))
hdr, units, body = raw[0], raw[1], raw[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]
for r in body:
    print("==", r[ix["Kernel Name"]].split("(")[0])
    print("   " + "  ".join("%s=%s%s" % (w.split("__")[-1].split(".")[0][:22], r[ix[w]], units[ix[w]] if units[ix[w]] != "%" else "%")
                            for w in want if w in ix))
rows = list(csv.reader(open("gpurun_out/%s_src.csv" % tag)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for sec in secs:
    if sec["name"] in seen:
        continue
    seen.add(sec["name"])
    h = sec["rows"][0]
    b = [r for r in sec["rows"][1:] if len(r) > 6]
    isrc, iex = h.index("Source"), h.index("Instructions Executed")
    tot = sum(int(r[iex] or 0) for r in b)
    hist = {}
    for r in b:
        op = r[isrc].strip().split()
        if not op:
            continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        hist[o] = hist.get(o, 0) + int(r[iex] or 0)
    print("== %s: %d SASS lines, %d warp instructions" % (sec["name"].split("(")[0], len(b), tot))
    print("   " + "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:top]))
