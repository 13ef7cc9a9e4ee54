"""Write-back / per-atom-sum passes on small structures at 100 000 frames: fused kernels against the split form (fit-only launch +
one elementwise sweep), switched with gcb_set_split_thresholds: python tools/split_small_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
lib = _lib.load()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for na in (50, 100, 200, 300, 440, 512, 600, 700, 800, 1000, 1250, 1500):
    r, s = synth.make_reference(na), synth.atom_sigmas(na)
    t = _lib.DeviceTraj(na, nf)
    t.synth_fill(r, s, nf)
    idx = np.arange(0, na, 2)
    res = {"natoms": na, "nframes": nf}
    for name, thr in (("fused", 0), ("split", 1 << 30)):
        t.set_split_thresholds(thr, thr)
        for mode in ("wb", "lovo", "wb_subset"):
            fn = {"wb": lambda: t.super_rmsd_async(r, write_back=True), "lovo": lambda: t.peratom_dev_sum(r, idx),
                  "wb_subset": lambda: t.super_rmsd_async(r, idx, idx, write_back=True)}[mode]
            fn(); fn(); t.sync(); t.timer_start()
            for _ in range(10):
                fn()
            res[f"{name}_{mode}_us"] = round(t.timer_stop() / 10 * 1e3, 1)
    t.set_split_thresholds(-1, -1)
    print(json.dumps(res), flush=True)
    t.close()
