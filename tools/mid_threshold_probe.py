"""Fit-only pass at 100 000 frames around the chunked-kernel / cluster-kernel boundary: python tools/mid_threshold_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
nf = 100000
for na in (1280, 1400, 1500, 1700, 2000, 2048):
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    idx = np.arange(0, na, 2)
    res = {"natoms": na, "lib": os.path.basename(_lib.LIB_PATH)}
    for mode in ("fit", "subset_fit"):
        fn = (lambda: t.super_rmsd_async(r)) if mode == "fit" else (lambda: t.super_rmsd_async(r, idx, idx))
        fn(); fn(); t.sync(); t.timer_start()
        for _ in range(10):
            fn()
        ms = t.timer_stop() / 10
        res[f"{mode}_us"] = round(ms * 1e3, 1); res[f"{mode}_GBs"] = round(24 * na * nf / ms / 1e6, 1)
    print(json.dumps(res), flush=True)
    t.close()
