"""Secondary entry points on a resident trajectory (one GPU): chem.RMSD without fit, Super with two different index lists
(gather path), Super with write-back, small frames, and the single-frame chem.Super latency of the host mirror.
    python tools/misc_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

out = {}
lib = _lib.load()


def timed(t, fn, reps=5):
    fn(); fn()
    t.sync()
    t.timer_start()
    for _ in range(reps):
        fn()
    return t.timer_stop() / reps


natoms, nframes = 10000, 100000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
gb = 24.0 * natoms * nframes / 1e6
ms = timed(t, lambda: t.rmsd(ref))
out["rmsd_nofit_100kx10k"] = {"ms_incl_d2h_of_rmsd": ms, "GBs": gb / ms}
ia = np.arange(0, natoms, 2)
templa = ref[ia]                       # a template that only holds the fitted atoms: lists (ia, 0..n-1) pair the same atoms
ib = np.arange(ia.size)
ms = timed(t, lambda: t.super_rmsd_async(templa, ia, ib))
out["super_two_lists_5000_of_10k_smaller_template"] = {"ms": ms, "frames_per_s": nframes / ms * 1e3, "GBs_full_frame": gb / ms}
ia2 = np.concatenate([ia, ia[:1]])     # one test atom twice: the gather kernels
ib2 = np.concatenate([ib, ib[1:2]])
ms = timed(t, lambda: t.super_rmsd_async(templa, ia2, ib2))
out["super_two_lists_with_a_repeated_test_atom_gather_kernel"] = {"ms": ms, "frames_per_s": nframes / ms * 1e3, "GBs_full_frame": gb / ms}
ms = timed(t, lambda: t.super_rmsd_async(ref, ia, ia))
out["super_subset_same_list_5000_of_10k"] = {"ms": ms, "frames_per_s": nframes / ms * 1e3, "GBs_full_frame": gb / ms}
ms = timed(t, lambda: t.super_rmsd_async(ref, write_back=True))
out["super_write_back_100kx10k"] = {"ms": ms, "frames_per_s": nframes / ms * 1e3, "GBs_48B": 2 * gb / ms}
t.close()
for na, nf in ((1500, 100000), (500, 200000), (100, 500000)):
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    tt = _lib.DeviceTraj(na, nf)
    tt.synth_fill(r, s, nf)
    ms = timed(tt, lambda: tt.super_rmsd_async(r))
    out["fit_%dx%d" % (nf, na)] = {"ms": ms, "frames_per_s": nf / ms * 1e3, "GBs": 24.0 * na * nf / ms / 1e6, "plan": tt.last_cluster_plan()}
    tt.close()
# a solvated system: 100 000 atoms per frame, the fit on a small subset (frames untouched: RMSD + rotation only)
na, nf = 100000, 20000
r, sg = synth.make_reference(na), synth.atom_sigmas(na)
tb = _lib.DeviceTraj(na, nf)
tb.synth_fill(r, sg, nf)
for n_sub in (300, 3000, 20000):
    sub = np.sort(np.random.default_rng(n_sub).choice(na, n_sub, replace=False))
    res = {}
    for name, var in (("auto", _lib.KERNEL_AUTO), ("generic", _lib.KERNEL_GENERIC)):
        tb.set_kernel_variant(var)
        ms = timed(tb, lambda: tb.super_rmsd_async(r, sub, sub))
        res[name + "_ms"] = ms
        res[name + "_frames_per_s"] = nf / ms * 1e3
    tb.set_kernel_variant(_lib.KERNEL_AUTO)
    out["solvated_100k_atoms_subset_%d" % n_sub] = res
tb.close()
# latency of one chem.Super on host matrices (batch of one through the host mirror)
a = synth.make_case(1500, 2)[1]
H.Super(a[1].copy(), a[0])
t0 = time.perf_counter()
for _ in range(200):
    H.Super(a[1].copy(), a[0])
out["chem_Super_single_frame_1500_atoms_us"] = (time.perf_counter() - t0) / 200 * 1e6
print(json.dumps(out))
