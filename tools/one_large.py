"""Profiling target: the split path on large frames.  python tools/one_large.py natoms nframes nidx"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
na, nf, n_sub = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
t = _lib.DeviceTraj(na, nf)
t.synth_fill(r, sg, nf)
sub = np.sort(np.random.default_rng(n_sub).choice(na, n_sub, replace=False)) if n_sub else None
for _ in range(2):
    t.super_rmsd_async(r, sub, sub, write_back=True)
    t.peratom_dev_sum(r, sub)
t.sync()
print("ok")
