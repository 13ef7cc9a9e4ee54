"""Profiling target: the elementwise sweeps of the split form.  python tools/one_split.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
# small structures: 300 atoms x 100 000 frames, write-back and per-atom sums
na, nf = 300, 100000
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
t = _lib.DeviceTraj(na, nf)
t.synth_fill(r, sg, nf)
for _ in range(2):
    t.super_rmsd_async(r, write_back=True)
    t.peratom_dev_sum(r, np.arange(0, na, 2))
t.sync()
t.close()
# a solvated system: 100 000 atoms x 4000 frames, the sums of 1000 candidate atoms
na, nf = 100000, 4000
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
t = _lib.DeviceTraj(na, nf)
t.synth_fill(r, sg, nf)
cand = np.sort(np.random.default_rng(5).choice(na, 1000, replace=False))
for _ in range(2):
    t.peratom_dev_sum_atoms(r, cand[:300], cand)
t.sync()
print("ok")
