"""Fit-only pass on a subset of a large frame (gather kernels): python tools/subset_fit_probe.py natoms nframes nidx [nidx ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth

na, nf = int(sys.argv[1]), int(sys.argv[2])
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
t = _lib.DeviceTraj(na, nf)
t.synth_fill(r, sg, nf)
out = {"natoms": na, "nframes": nf, "lib": os.path.basename(_lib.LIB_PATH)}
for n_sub in [int(a) for a in sys.argv[3:]]:
    sub = np.sort(np.random.default_rng(n_sub).choice(na, n_sub, replace=False))
    t.super_rmsd_async(r, sub, sub)
    t.sync()
    t.timer_start()
    for _ in range(5):
        t.super_rmsd_async(r, sub, sub)
    out["subset_%d_ms" % n_sub] = t.timer_stop() / 5
print(json.dumps(out))
