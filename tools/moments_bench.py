"""Per-frame centre of mass + moment tensor over a resident trajectory: python tools/moments_bench.py [natoms] [nframes]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from gochem_b200 import _lib, synth
natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
mass = np.random.default_rng(0).uniform(1, 32, natoms)
t.moments(mass)
t0 = time.perf_counter()
cen, mom = t.moments(mass)
dt = time.perf_counter() - t0
fr = t.download(5, 1)[0]
oc, om = oracle.center_and_moment(fr, mass)
print(json.dumps({"natoms": natoms, "nframes": nframes, "seconds_incl_d2h": dt, "frames_per_s": nframes / dt,
                  "GBs": 24.0 * natoms * nframes / dt / 1e9, "max_rel_err_vs_oracle": float(np.abs(mom[5] - om).max() / np.abs(om).max())}))
