"""auto (warp-per-frame kernels where the dispatch picks them) vs the cluster kernel, fit-only / write-back / per-atom sums, 100 000 frames"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from gochem_b200 import _lib, synth
nf = 100000
for na in (300, 448, 512, 640, 800, 900, 1000, 1152, 1280, 1408, 1536, 1664, 2000, 3000, 5000):
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    idx = np.arange(0, na, 2)
    res = {"natoms": na}
    for var, name in ((_lib.KERNEL_AUTO, "auto"), (_lib.KERNEL_CLUSTER, "cluster")):
        t.set_kernel_variant(var)
        for mode in ("fit", "wb", "lovo"):
            fn = {"fit": lambda: t.super_rmsd_async(r), "wb": lambda: t.super_rmsd_async(r, write_back=True), "lovo": lambda: t.peratom_dev_sum(r, idx)}[mode]
            try:
                fn(); fn(); t.sync(); t.timer_start()
                for _ in range(10):
                    fn()
                ms = t.timer_stop() / 10
                res[f"{name}_{mode}_us"] = round(ms * 1e3, 1)
            except _lib.GcbError as e:
                res[f"{name}_{mode}_us"] = None
    print(json.dumps(res), flush=True)
    t.close()
