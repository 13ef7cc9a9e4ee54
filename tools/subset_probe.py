"""Probe: fit-only passes with and without a subset mask, alternating, with the SM / memory clocks sampled beside them."""
import os, sys, json, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
natoms, nframes = 10000, 100000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
ia = np.arange(0, natoms, 2)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "50"],
                       stdout=open("gpurun_out/s3u/clocks.csv", "w"))
def timed(fn, reps=5):
    fn(); fn(); t.sync(); t.timer_start()
    for _ in range(reps): fn()
    return round(t.timer_stop() / reps, 4)
r = []
for rep in range(6):
    r.append(("masked", timed(lambda: t.super_rmsd_async(ref, ia, ia))))
    r.append(("all", timed(lambda: t.super_rmsd_async(ref))))
ib = np.arange(0, natoms, 2) + 1
for rep in range(3):
    r.append(("masked_odd", timed(lambda: t.super_rmsd_async(ref, ib, ib))))
    r.append(("masked_even", timed(lambda: t.super_rmsd_async(ref, ia, ia))))
fl = np.frombuffer(b"", dtype=np.uint8)
time.sleep(0.2)
smi.terminate()
print(json.dumps(r))
