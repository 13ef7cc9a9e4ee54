"""Profiling target: a few passes over small structures.  python tools/one_small.py natoms nframes [reps] [fit|wb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gochem_b200 import _lib, synth
natoms, nframes = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wb = len(sys.argv) > 4 and sys.argv[4] == "wb"
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
t.super_rmsd_async(ref, write_back=wb)
t.sync()
t.timer_start()
for _ in range(reps):
    t.super_rmsd_async(ref, write_back=wb)
ms = t.timer_stop() / reps
print("ms per pass", ms, "GB/s", (48 if wb else 24) * natoms * nframes / ms / 1e6, "Mframes/s", nframes / ms / 1e3)
