"""BASELINE configs[4] (per-GPU share): frames streamed from C-owned pinned host memory through the staging path
(gcb_traj_upload_async, two copy streams) into fused superposition + RMSD.
    python tools/stream_bench.py [nframes] [natoms] [f32soa|f32aos|f64aos]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
natoms = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
fmt = sys.argv[3] if len(sys.argv) > 3 else "f32soa"
dev = int(os.environ.get("LOCAL_RANK", "0"))
layout = {"f32soa": _lib.F32_SOA, "f32aos": _lib.F32_AOS, "f64aos": _lib.F64_AOS}[fmt]
dt = np.float64 if fmt == "f64aos" else np.float32
fbytes = natoms * 3 * np.dtype(dt).itemsize
chunk = max(1, (64 << 20) // fbytes)
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
# what a reader would decode straight into the pinned ring: one chunk of real frames, reused for every chunk
src = _lib.DeviceTraj(natoms, chunk, dev=dev)
src.synth_fill(ref, sig, chunk, through_f32=(fmt != "f64aos"))
host = src.download()
src.close()
if fmt == "f32soa":
    host = np.ascontiguousarray(host.astype(np.float32).transpose(0, 2, 1))
elif fmt == "f32aos":
    host = host.astype(np.float32)
ring = [_lib.PinnedBuffer(chunk * fbytes) for _ in range(2)]
views = [r.view(dt, host.shape) for r in ring]
for v in views:
    v[...] = host
t = _lib.DeviceTraj(natoms, 2 * chunk, dev=dev)
nchunks = (nframes + chunk - 1) // chunk


def run(compute: bool):
    t.sync()
    t0 = time.perf_counter()
    done = 0
    for c in range(nchunks):
        n = min(chunk, nframes - done)
        s = c & 1
        t.wait(s)                                     # the reader may refill ring slot s from here on
        t.upload_async(views[s][:n], n, s * chunk, layout, 1.0, s)
        if compute:
            t.super_rmsd_async(ref, first=s * chunk, nframes=n)
        done += n
    t.sync()
    return time.perf_counter() - t0


run(True)
copy_only = run(False)
total = run(True)
rm = t.fetch(0, chunk, want=("rmsd",))["rmsd"]
print(json.dumps({"format": fmt, "nframes": nframes, "natoms": natoms, "chunk_frames": chunk, "seconds": total,
                  "frames_per_s": nframes / total, "h2d_GBs": nframes * fbytes / total / 1e9, "copy_only_seconds": copy_only,
                  "overlap_efficiency": copy_only / total, "rmsd_mean_last_chunk": float(rm.mean())}), flush=True)
