"""BASELINE configs[0]: chem.Super + chem.RMSD of a 1000-frame x 1500-atom trajectory against frame 0, end to end from host
float64 matrices (upload, fused pass with the superposed frames written back, download of frames and per-frame results), beside
the CPU port of the reference and its as-written cost model (N x N identity product per frame) on this host's cores.
    python tools/config1_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from gochem_b200 import _lib, synth

natoms, nframes = 1500, 1000
ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
t = _lib.DeviceTraj(natoms, nframes)


def run():
    t.upload(frames)
    out = t.super_rmsd(ref, write_back=True)
    return out, t.download()


run()
times = []
for _ in range(10):
    t0 = time.perf_counter()
    out, moved = run()
    times.append(time.perf_counter() - t0)
t.upload(frames)
t.super_rmsd_async(ref, write_back=True); t.sync()
t.upload(frames)
t.timer_start()
t.super_rmsd_async(ref, write_back=True)
kernel_ms = t.timer_stop()
cores = os.cpu_count() or 1
work = frames.copy()
t0 = time.perf_counter()
rm, rot, t1, t2 = oracle.super_rmsd_traj(work, ref, None, write_back=True, nthreads=cores)
port_s = time.perf_counter() - t0
sample = frames[:4 * cores].copy()
t0 = time.perf_counter()
oracle.super_rmsd_traj(sample, ref, None, write_back=True, as_written=True, nthreads=cores, details=False)
aw_s = (time.perf_counter() - t0) * nframes / sample.shape[0]
print(json.dumps({"workload": "chem.Super + chem.RMSD, 1000 frames x 1500 atoms against frame 0 (BASELINE configs[0])",
                  "gpu_end_to_end_ms_min": min(times) * 1e3, "gpu_end_to_end_ms_median": sorted(times)[5] * 1e3,
                  "gpu_end_to_end_frames_per_s": nframes / min(times), "kernel_only_ms": kernel_ms,
                  "cpu_port_ms": port_s * 1e3, "cpu_as_written_ms_extrapolated": aw_s * 1e3, "cpu_threads": cores,
                  "max_abs_rmsd_diff": float(np.abs(out["rmsd"] - rm).max()), "max_abs_rot_diff": float(np.abs(out["rot"] - rot).max()),
                  "max_abs_coord_diff": float(np.abs(moved - work).max()), "rmsd_frame0": float(out["rmsd"][0])}))
t.close()
