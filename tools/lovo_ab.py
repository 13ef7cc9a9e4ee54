"""align.LOVO on resident frames: the device-side bookkeeping (gcb_lovo_resident) against the host loop (GCH_LOVO_HOST=1), same
process, same frames, alternating.  python tools/lovo_ab.py [natoms] [nframes]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
t.sync()
best = {"device": 1e9, "host": 1e9}
for rnd in range(6):
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["GCH_LOVO_HOST"] = "1"
        else:
            os.environ.pop("GCH_LOVO_HOST", None)
        for _ in range(3):
            t0 = time.perf_counter()
            got = H.LOVOResident(t, ref, cpus=8)
            dt = time.perf_counter() - t0
            if rnd > 0:
                best[mode] = min(best[mode], dt)
print(json.dumps({"natoms": natoms, "nframes": nframes, "iterations": got["Iterations"], "N": got["N"],
                  "lovo_ms_device_loop": best["device"] * 1e3, "lovo_ms_host_loop": best["host"] * 1e3}))
