"""BASELINE configs[4] through the reference-facing API: align.RMSDTraj on a DCD FILE, streamed through the device window
(gochem::align::RMSDTrajStreamed): file (page cache) -> DCD reader -> C-owned pinned float32 planes -> device -> fused pass.
    python tools/dcd_stream_bench.py [natoms] [nframes]"""
import json, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes, through_f32=True)
frames = t.download()
t.close()
path = os.path.join(tempfile.gettempdir(), "gcb_stream_bench.dcd")
H.dcd_write(path, frames)
size = os.path.getsize(path)
idx = np.arange(natoms)
cpus = 8
H.RMSDTrajStreamed(path, ref, idx, cpus, 0)           # warm-up (page cache, allocations); window 0 = 64 MB default
t0 = time.perf_counter()
got, n = H.RMSDTrajStreamed(path, ref, idx, cpus, 0)
dt = time.perf_counter() - t0
t0 = time.perf_counter()
res, n2 = H.RMSDTraj(path, ref, idx, cpus)             # resident path: whole file to host memory, then upload, then one pass
dt_res = time.perf_counter() - t0
t0 = time.perf_counter()
back, nf, na = H.dcd_read(path, 8, natoms)             # reader alone would be: time a raw read of the file
with open(path, "rb") as f:
    while f.read(64 << 20):
        pass
dt_read = time.perf_counter() - t0
t0 = time.perf_counter()
lv = H.LOVO(path, ref, cpus)                           # align.LOVO from the file: load once, iterate on the resident frames
dt_lovo = time.perf_counter() - t0
t0 = time.perf_counter()
ls = H.LOVOStreamed(path, ref, cpus, 0)                # the file re-read on every iteration (lovo.go:222), bounded memory
dt_lovo_s = time.perf_counter() - t0
os.remove(path)
print(json.dumps({"lovo_resident_seconds": dt_lovo, "lovo_streamed_seconds": dt_lovo_s, "lovo_iterations": lv["Iterations"],
                  "lovo_same_core": bool(list(lv["Natoms"]) == list(ls["Natoms"])), "lovo_N": lv["N"],
                  "natoms": natoms, "nframes": n, "file_GB": size / 1e9, "streamed_seconds": dt, "streamed_frames_per_s": n / dt,
                  "streamed_file_GBs": size / dt / 1e9, "resident_seconds": dt_res, "resident_frames_per_s": n2 / dt_res,
                  "plain_file_read_seconds": dt_read, "max_abs_diff_streamed_vs_resident": float(np.abs(got - res).max())}))
