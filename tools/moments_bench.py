"""Per-frame centre of mass + moment tensor over a resident trajectory: python tools/moments_bench.py [natoms] [nframes]
Kernel time from CUDA events on the library's stream (gcb_timer_*), 24 B per atom per frame."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from gochem_b200 import _lib, synth
natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = 10
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
mass = np.random.default_rng(0).uniform(1, 32, natoms)
for _ in range(3):
    t.moments_async(mass)
t.sync()
t.timer_start()
for _ in range(reps):
    t.moments_async(mass)
ms = t.timer_stop() / reps
t0 = time.perf_counter()
cen, mom = t.moments(mass)
wall = time.perf_counter() - t0
err = 0.0
for f in (0, 5, nframes - 1):
    fr = t.download(f, 1)[0]
    oc, om = oracle.center_and_moment(fr, mass)
    err = max(err, float(np.abs(mom[f] - om).max() / np.abs(om).max()), float(np.abs(cen[f] - oc).max()))
print(json.dumps({"natoms": natoms, "nframes": nframes, "kernel_ms": ms, "frames_per_s": nframes / ms * 1e3,
                  "GBs": 24.0 * natoms * nframes / ms / 1e6, "wall_s_incl_d2h": wall, "max_err_vs_oracle": err,
                  "note": "gcb_moments_async: moments_stream_kernel (TMA-chunked persistent stream) incl. the 80 KB mass-plane upload per call"}))
