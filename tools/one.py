"""Run one kernel configuration a few times (profiling target): python tools/one.py natoms nframes mode C S D [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth

natoms, nframes, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
c, s, d = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
lib = _lib.load()
ref = synth.make_reference(natoms)
sig = synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes, frame0_is_ref=True)
if c < 0:
    t.set_kernel_variant(_lib.KERNEL_GENERIC)
else:
    t.set_kernel_variant(_lib.KERNEL_AUTO)
    t.set_cluster_tuning(c, s, d, int(os.environ.get('GCB_RES', '-1')))
idx = np.arange(0, natoms, 2)
t.sync()
t.timer_start()
for _ in range(reps):
    if mode == "fit":
        t.super_rmsd_async(ref)
    elif mode == "wb":
        t.super_rmsd_async(ref, write_back=True)
    else:
        t.peratom_dev_sum(ref, idx)
ms = t.timer_stop() / reps
print("ms per pass", ms, "GB/s", (48 if mode == "wb" else 24) * natoms * nframes / ms / 1e6)
