"""Build a variant of libgochem_b200.so with extra nvcc flags (tuning / trace builds), next to the product library:
    python tools/build_variant.py trace -DGCB_TRACE
    python tools/build_variant.py sf --only super_cluster.cu -DGCB_SOLVER_FIRST     (other objects come from the product build)
-> gochem_b200/lib/libgochem_b200_<suffix>.so; use it with GCB_LIB=<path> (gochem_b200/_lib.py)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gochem_b200 import build as b

args = sys.argv[1:]
suffix = args.pop(0)
only = None
if args and args[0] == "--only":
    only = args[1].split(",")
    args = args[2:]
flags = args
b.build()
obj_dir = os.path.join(b.HERE, "build_" + suffix)
os.makedirs(obj_dir, exist_ok=True)
procs, objs = [], []
for src in b.SOURCES:
    if only is not None and src not in only:
        objs.append(os.path.join(b.HERE, "build", src.replace(".cu", ".o")))
        continue
    o = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen([b.NVCC, *b.ARCH, *b.FLAGS, *flags, "-c", os.path.join(b.CSRC, src), "-o", o]))
if any(p.wait() for p in procs):
    raise SystemExit("nvcc failed")
out = os.path.join(b.LIB_DIR, "libgochem_b200_%s.so" % suffix)
subprocess.check_call([b.NVCC, *b.ARCH, "-shared", "-o", out, *objs, "-ldl"])
print(out)
