"""Fit-only pass: warp-per-frame kernels (KERNEL_AUTO below 2048 atoms) against the cluster kernel, by structure size.
    python tools/small_sweep.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
lib = _lib.load()
for na in (50, 100, 200, 300, 512, 600, 800, 1000, 1250, 1500, 2000):
    nf = int(min(2_000_000, 3.6e9 / (24 * na)))
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    res = {"natoms": na, "nframes": nf}
    for name, var in (("auto", _lib.KERNEL_AUTO), ("cluster", _lib.KERNEL_CLUSTER)):
        t.set_kernel_variant(var)
        for mode in ("fit", "wb"):
            fn = (lambda: t.super_rmsd_async(r)) if mode == "fit" else (lambda: t.super_rmsd_async(r, write_back=True))
            fn(); fn(); t.sync(); t.timer_start()
            for _ in range(5):
                fn()
            ms = t.timer_stop() / 5
            res[f"{name}_{mode}_GBs"] = round((24 if mode == "fit" else 48) * na * nf / ms / 1e6, 1)
            res[f"{name}_{mode}_Mframes_s"] = round(nf / ms / 1e3, 1)
    t.set_kernel_variant(_lib.KERNEL_AUTO)
    print(json.dumps(res), flush=True)
    t.close()
