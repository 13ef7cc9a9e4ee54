"""Opcode histogram per kernel of libgochem_b200.so (cuobjdump -sass): the evidence for which hardware paths the kernels use
(UBLKCP = 1-D TMA bulk copy, UTMALDG = tiled TMA load, SYNCS = mbarrier, STAS = st.async over DSMEM, DMMA = fp64 tensor core,
LDGSTS = cp.async).  python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gochem_b200", "lib", "libgochem_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|gcb::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
marks = ["UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "STAS", "DMMA", "LDGSTS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "RED",
         "SHFL", "BAR", "UCGABAR_ARV", "MEMBAR", "LDL", "STL"]
total = collections.Counter()
print("library:", os.path.relpath(lib, ROOT))
print("%-78s %7s  %s" % ("kernel", "instrs", "  ".join(marks)))
for k, h in hist.items():
    n = sum(h.values())
    total.update(h)
    print("%-78s %7d  %s" % (k[:78], n, "  ".join("%*d" % (len(m), h.get(m, 0)) for m in marks)))
print("%-78s %7d  %s" % ("ALL KERNELS", sum(total.values()), "  ".join("%*d" % (len(m), total.get(m, 0)) for m in marks)))
print("\nnot present anywhere (expected: fp64 has no tcgen05 kind, so no UTC*MMA / LDTM):",
      [m for m in ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "HGMMA"] if not any(k.startswith(m) for k in total)])
