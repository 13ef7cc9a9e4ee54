"""Fit-only and write-back passes over 100 000 frames (a typical trajectory length) of small structures: python tools/frames100k_sweep.py [nframes]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
lib = _lib.load()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for na in (50, 100, 200, 300, 440, 512, 700, 1000, 1250):
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    res = {"natoms": na, "nframes": nf, "lib": os.path.basename(_lib.LIB_PATH)}
    idx = np.arange(0, na, 2)
    for mode in ("fit", "wb", "lovo"):
        fn = {"fit": lambda: t.super_rmsd_async(r), "wb": lambda: t.super_rmsd_async(r, write_back=True), "lovo": lambda: t.peratom_dev_sum(r, idx)}[mode]
        fn(); fn(); t.sync(); t.timer_start()
        for _ in range(10):
            fn()
        ms = t.timer_stop() / 10
        res[f"{mode}_us"] = round(ms * 1e3, 1)
        res[f"{mode}_GBs"] = round((48 if mode == "wb" else 24) * na * nf / ms / 1e6, 1)
    print(json.dumps(res), flush=True)
    t.close()
