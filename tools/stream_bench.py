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
if os.environ.get("GCB_NUMA_BIND", "1") == "1":
    rc = _lib.load().gcb_bind_host_thread(dev)     # pinned ring NUMA-local to this GPU's PCIe link
    if rc:
        print("bind_host_thread:", _lib.last_error(), file=sys.stderr)
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


world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:   # all ranks of one node stream at the same time: each has its own ring, copy streams and PCIe link
    import torch.distributed as dist

    dist.init_process_group("gloo")
    barrier = dist.barrier
else:
    barrier = lambda: None
run(True)
barrier()
copy_only = run(False)
barrier()
total = run(True)
rm = t.fetch(0, chunk, want=("rmsd",))["rmsd"]
res = {"format": fmt, "nframes": nframes, "natoms": natoms, "chunk_frames": chunk, "seconds": total,
       "frames_per_s": nframes / total, "h2d_GBs": nframes * fbytes / total / 1e9, "copy_only_seconds": copy_only,
       "overlap_efficiency": copy_only / total, "rmsd_mean_last_chunk": float(rm.mean())}
if world > 1:
    g = [None] * world
    dist.all_gather_object(g, res)
    if rank == 0:
        slowest = max(x["seconds"] for x in g)
        print(json.dumps({"format": fmt, "world": world, "frames_per_rank": nframes, "natoms": natoms, "seconds_max_over_ranks": slowest,
                          "frames_per_s_aggregate": world * nframes / slowest, "h2d_GBs_aggregate": world * nframes * fbytes / slowest / 1e9,
                          "per_rank_h2d_GBs": [round(x["h2d_GBs"], 2) for x in g],
                          "per_rank_overlap": [round(x["overlap_efficiency"], 3) for x in g]}), flush=True)
    dist.destroy_process_group()
else:
    print(json.dumps(res), flush=True)
