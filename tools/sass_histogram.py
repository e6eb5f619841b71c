"""Static SASS opcode histogram per kernel of the built library (cuobjdump -sass): the mnemonics that prove the
Blackwell path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit) and the
instruction mix of the CUDA-core kernels.  usage: python tools/sass_histogram.py [lib.so] > profiles/sass_histogram_rN.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "densebox_b200/csrc/libdensebox_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = kern.replace("void dbx::", "").replace("dbx::", "")
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLKCP")) else m.group(1)] += 1
print("cuobjdump -sass %s   (ELF arch: %s)" % (lib, ", ".join(sorted(arch))))
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA")
for k, h in hist.items():
    tot = sum(h.values())
    key = {n: c for n, c in h.items() if n.startswith(KEY)}
    top = ", ".join("%s %d" % (n, c) for n, c in h.most_common(10))
    print("\n%s: %d SASS instructions" % (k, tot))
    if key:
        print("   tensor/TMA: " + ", ".join("%s %d" % (n, c) for n, c in sorted(key.items())))
    print("   top: " + top)
