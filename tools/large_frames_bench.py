"""Solvated systems: frames too large for the cluster kernel (100 000 atoms), passes that visit every atom twice.
Super with write-back and align.RMSDTraj's per-atom sums, fit on a small subset / on all atoms; the split path of the
library (fit launch + elementwise sweep, transform.cu) against the one-CTA-per-frame kernel that re-reads the frame.
    python tools/large_frames_bench.py [natoms nframes]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth

lib = _lib.load()
na, nf = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 20000)
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
t = _lib.DeviceTraj(na, nf)
t.synth_fill(r, sg, nf)


def timed(fn, reps=3):
    fn()
    t.sync()
    t.timer_start()
    for _ in range(reps):
        fn()
    return t.timer_stop() / reps


out = {"natoms": na, "nframes": nf}
for n_sub in (300, 3000, 0):
    sub = np.sort(np.random.default_rng(n_sub).choice(na, n_sub, replace=False)) if n_sub else None
    res = {}
    for name, var in (("auto", _lib.KERNEL_AUTO), ("generic", _lib.KERNEL_GENERIC)):
        t.set_kernel_variant(var)
        ms = timed(lambda: t.super_rmsd_async(r, sub, sub, write_back=True))
        res[name + "_write_back_ms"] = ms
        res[name + "_write_back_GBs_48B"] = 48.0 * na * nf / ms / 1e6
        ms = timed(lambda: t.peratom_dev_sum(r, sub))
        res[name + "_per_atom_sums_ms"] = ms
        res[name + "_per_atom_sums_GBs_24B"] = 24.0 * na * nf / ms / 1e6
    t.set_kernel_variant(_lib.KERNEL_AUTO)
    out["fit_on_%s" % (n_sub or "all")] = res
# align.LOVO on such a system only needs the sums of its candidates (the CA atoms): 1000 atoms asked for, fit on 300 of them
cand = np.sort(np.random.default_rng(5).choice(na, 1000, replace=False))
fit = cand[:300]
ms = timed(lambda: t.peratom_dev_sum_atoms(r, fit, cand))
out["candidates_only_1000_atoms_fit_on_300"] = {"per_atom_sums_ms": ms, "full_frame_entry_point_ms": timed(lambda: t.peratom_dev_sum(r, fit))}
t.close()
print(json.dumps(out))
