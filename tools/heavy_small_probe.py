"""Passes that visit every atom twice (write-back, per-atom sums) on small structures, built-in dispatch: python tools/heavy_small_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
lib = _lib.load()
for na in (200, 256, 300, 400, 512):
    nf = int(min(2_000_000, 3.6e9 / (24 * na)))
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    idx = np.arange(0, na, 2)
    res = {"natoms": na, "nframes": nf, "lib": os.path.basename(_lib.LIB_PATH)}
    for mode in ("wb", "lovo"):
        fn = (lambda: t.super_rmsd_async(r, write_back=True)) if mode == "wb" else (lambda: t.peratom_dev_sum(r, idx))
        fn(); fn(); t.sync(); t.timer_start()
        for _ in range(5):
            fn()
        ms = t.timer_stop() / 5
        res[f"{mode}_ms"] = round(ms, 4)
        res[f"{mode}_GBs"] = round((48 if mode == "wb" else 24) * na * nf / ms / 1e6, 1)
    print(json.dumps(res), flush=True)
    t.close()
