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
_lib.check(lib.gcb_set_cluster_tuning(c, s, d, int(os.environ.get('GCB_RES', '-1'))))
for _ in range(2):
    t.super_rmsd_async(ref)
t.sync()
buf = np.zeros((8, 512), dtype=np.int64)
assert lib.gcb_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
names = ["S:red_ready", "S:pushed", "S:parts_ready", "S:solved", "C:iter_start", "C:full_ready", "C:solved_seen", "C:red_arrive"]
sl = slice(40, 120)
b = buf[:, sl].astype(np.float64)
print("per-frame period (C:red_arrive):", np.diff(b[7]).mean())
print("warp0 red_arrive -> pushed (last warp sums + st.async)", (b[1] - b[7]).mean())
print("pushed -> S:parts_ready    ", (b[2] - b[1]).mean())
print("S:parts_ready -> S:solved  ", (b[3] - b[2]).mean())
print("S:solved -> C:solved_seen  ", (b[6] - b[3]).mean())
print("total red_arrive -> solved_seen", (b[6] - b[7]).mean())
print("C:iter_start -> full_ready (wait for data)", (b[5] - b[4]).mean())
print("C:full_ready -> red_arrive (pass1 [+pass2 incl. solved wait] + reduce)", (b[7] - b[5]).mean())
