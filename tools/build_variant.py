"""Build a variant of libgochem_b200.so with extra nvcc flags (tuning / trace builds), next to the product library:
    python tools/build_variant.py trace -DGCB_TRACE
-> gochem_b200/lib/libgochem_b200_trace.so; use it with GCB_LIB=<path> (gochem_b200/_lib.py)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gochem_b200 import build as b

suffix, flags = sys.argv[1], sys.argv[2:]
obj_dir = os.path.join(b.HERE, "build_" + suffix)
os.makedirs(obj_dir, exist_ok=True)
procs, objs = [], []
for src in b.SOURCES:
    o = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen([b.NVCC, *b.ARCH, *b.FLAGS, *flags, "-c", os.path.join(b.CSRC, src), "-o", o]))
if any(p.wait() for p in procs):
    raise SystemExit("nvcc failed")
out = os.path.join(b.LIB_DIR, "libgochem_b200_%s.so" % suffix)
subprocess.check_call([b.NVCC, *b.ARCH, "-shared", "-o", out, *objs, "-ldl"])
print(out)
