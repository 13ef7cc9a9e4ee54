"""Cycle trace of CTA 0 of the cluster kernel (needs a library built with GCB_TRACE=1)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth

natoms, nframes = int(sys.argv[1]), int(sys.argv[2])
c, s, d = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
lib = _lib.load()
ref = synth.make_reference(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, synth.atom_sigmas(natoms), nframes, frame0_is_ref=True)
t.set_cluster_tuning(c, s, d, int(os.environ.get('GCB_RES', '-1')))
mode = sys.argv[6] if len(sys.argv) > 6 else "fit"
idx = np.arange(0, natoms, 2)
for _ in range(2):
    if mode == "fit":
        t.super_rmsd_async(ref)
    elif mode == "wb":
        t.super_rmsd_async(ref, write_back=True)
    else:
        t.peratom_dev_sum(ref, idx)
t.sync()
buf = np.zeros((8, 512), dtype=np.int64)
assert lib.gcb_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
plan = t.last_cluster_plan()
D = plan["depth"]
print("plan", plan)
sl = slice(40, 120)
b = buf.astype(np.float64)
per = np.diff(b[7, sl]).mean()
print("per-frame period (C:red_arrive)            : %.0f" % per)
print("C: wait for the stage (full)               : %.0f" % (b[5, sl] - b[4, sl]).mean())
print("C: pass 1 + warp reduce (full->red_arrive) : %.0f" % (b[7, sl] - b[5, sl]).mean())
print("   red_arrive(warp0) -> pushed (last warp) : %.0f" % (b[1, sl] - b[7, sl]).mean())
print("S: pushed -> partials of all CTAs landed   : %.0f" % (b[2, sl] - b[1, sl]).mean())
print("S: rank-ordered sum + 3x3 solve            : %.0f" % (b[3, sl] - b[2, sl]).mean())
print("   solved -> consumer resumes pass 2       : %.0f" % (b[6, sl] - b[3, sl]).mean())
print("   deliver(j) -> pass 2 of j starts        : %.0f" % (b[6, sl] - b[7, sl]).mean())
nxt = slice(sl.start + 1, sl.stop + 1)
print("C: pass 2 (solved_seen(j-D) -> iter_start(next)) : %.0f" % (b[4, slice(sl.start + D + 1, sl.stop + D + 1)] - b[6, sl]).mean())
