"""Sweep the cluster kernel's shape on one GPU: python tools/tune.py [natoms] [nframes] [mode]"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mode = sys.argv[3] if len(sys.argv) > 3 else "fit"
lib = _lib.load()
ref = synth.make_reference(natoms)
sig = synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes, frame0_is_ref=True)
idx = np.arange(0, natoms, 2)


def run(label):
    def step():
        if mode == "fit":
            t.super_rmsd_async(ref)
        elif mode == "wb":
            t.super_rmsd_async(ref, write_back=True)
        elif mode == "lovo":
            t.peratom_dev_sum(ref, idx)
    for _ in range(2):
        step()
    t.sync()
    t.timer_start()
    n = 5
    for _ in range(n):
        step()
    ms = t.timer_stop() / n
    bpa = 48 if mode == "wb" else 24
    gbs = bpa * natoms * nframes / ms / 1e6
    print(json.dumps({"cfg": label, "ms": round(ms, 4), "GBs": round(gbs, 1), "Mframes_s": round(nframes / ms / 1e3, 3)}), flush=True)


t.set_kernel_variant(_lib.KERNEL_GENERIC)
run("generic")
t.set_kernel_variant(_lib.KERNEL_CLUSTER)
cfgs = []
for res in (0, 1):
    for c in (1, 2, 3, 4, 5, 6, 7, 8):
        for st in ((0, 2, 3) if res == 0 else (0,)):
            for d in ((2, 4, 6) if res == 0 else (1, 2, 3)):
                cfgs.append((c, st, d, res))
if len(sys.argv) > 4:
    cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[4:]]
for c, st, d, res in cfgs:
    try:
        t.set_cluster_tuning(c, st, d, res)
        run(f"C{c} S{st} D{d} R{res}")
        print("   plan", t.last_cluster_plan(), flush=True)
    except _lib.GcbError as e:
        print(json.dumps({"cfg": f"C{c} S{st} D{d} R{res}", "err": e.msg[:80]}), flush=True)
