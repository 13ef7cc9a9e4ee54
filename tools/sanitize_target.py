"""Small instances of every kernel family, for compute-sanitizer (tools/sanitize.sh).  Sizes are tiny on purpose: the tools
slow kernels down 10-100x.  Answers are checked against the oracle so a run also proves the instrumented kernels still agree.
    python tools/sanitize_target.py [family ...]      families: small mid cluster generic gather split misc allpairs moments rdf"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from gochem_b200 import _lib, synth

lib = _lib.load()
want = set(sys.argv[1:]) or {"small", "mid", "cluster", "generic", "gather", "split", "misc", "allpairs", "moments", "rdf"}


def check_super(natoms, nframes, idx=None, variant=_lib.KERNEL_AUTO, tuning=None, label=""):
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    t = _lib.DeviceTraj(natoms, nframes)
    t.upload(frames)
    t.set_kernel_variant(variant)
    if tuning:
        t.set_cluster_tuning(*tuning)
    try:
        out = t.super_rmsd(ref, idx, idx)
        rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, idx)
        assert np.abs(out["rmsd"] - rm).max() < 1e-9 and np.abs(out["rot"] - rot).max() < 1e-10, label
        sel = idx if idx is not None else np.arange(natoms)
        dev = t.peratom_dev_sum(ref, idx)
        cpus = 1
        want_dev, total = oracle.rmsd_traj(frames, ref, sel, cpus=cpus)
        assert np.abs(dev / nframes - want_dev).max() < 1e-10, label
        out = t.super_rmsd(ref, idx, idx, write_back=True)
        work = frames.copy()
        oracle.super_rmsd_traj(work, ref, idx, write_back=True)
        assert np.abs(t.download() - work).max() < 1e-9, label
    finally:
        t.set_cluster_tuning(0, 0, 0, -1)
        t.set_kernel_variant(_lib.KERNEL_AUTO)
        t.close()
    print("ok", label or f"{natoms} atoms x {nframes} frames", flush=True)


if "small" in want:
    for n in (3, 50, 257, 512):
        check_super(n, 70, label=f"small {n}")
    check_super(200, 70, idx=np.arange(0, 200, 3), label="small subset")
if "mid" in want:
    check_super(1000, 40, label="mid 1000")
    check_super(1280, 33, idx=np.arange(0, 1280, 2), label="mid subset")
if "cluster" in want:
    check_super(1500, 40, label="cluster C1")
    check_super(5000, 12, idx=np.arange(0, 5000, 2), label="cluster C2 subset")
    check_super(3000, 20, variant=_lib.KERNEL_CLUSTER, tuning=(3, 0, 2, 1), label="cluster C3 resident D2")
    check_super(3000, 20, variant=_lib.KERNEL_CLUSTER, tuning=(4, 2, 6, 0), label="cluster C4 L2 pass 2")
if "generic" in want:
    check_super(700, 20, variant=_lib.KERNEL_GENERIC, label="generic")
    check_super(700, 20, idx=np.array([5, 5, 9, 300, 301, 302, 650]), variant=_lib.KERNEL_GENERIC, label="generic repeated index")
if "gather" in want:
    check_super(9000, 6, idx=np.sort(np.random.default_rng(1).choice(9000, 300, replace=False)), label="gather small")
    check_super(33000, 4, idx=np.sort(np.random.default_rng(2).choice(33000, 3000, replace=False)), label="gather chunked + split")
if "split" in want:
    check_super(24000, 3, label="split 24000 all atoms")
    check_super(30000, 3, idx=np.sort(np.random.default_rng(3).choice(30000, 200, replace=False)), label="split 30000 subset")
if "misc" in want:
    ref, frames = synth.make_case(400, 9)
    t = _lib.DeviceTraj(400, 9)
    t.upload(frames.astype(np.float32), layout=_lib.F32_AOS, scale=1.0)
    got = t.rmsd(ref)
    for f in range(9):
        assert abs(got[f] - oracle.rmsd(frames[f].astype(np.float32).astype(np.float64), ref)) < 1e-9
    t.close()
    print("ok misc (f32 upload, rmsd without fit)", flush=True)
if "allpairs" in want:
    ref, frames = synth.make_case(333, 130)
    t = _lib.DeviceTraj(333, 130)
    t.upload(frames)
    d, _, _ = t.allpairs_rmsd()
    x = frames
    i, j = 3, 97
    assert abs(d[i, j] - np.sqrt(((x[i] - x[j]) ** 2).sum() / 333)) < 1e-9
    d2, _, _ = t.allpairs_rmsd(fit=True)
    assert np.isfinite(d2).all() and (d2 <= d + 1e-9).all()
    t.close()
    print("ok allpairs", flush=True)
if "moments" in want:
    ref, frames = synth.make_case(2500, 7)
    t = _lib.DeviceTraj(2500, 7)
    t.upload(frames)
    mass = np.linspace(1, 16, 2500)
    cen, mom = t.moments(mass)
    c0 = (frames[0] * mass[:, None]).sum(0) / mass.sum()
    assert np.abs(cen[0] - c0).max() < 1e-9
    t.close()
    print("ok moments", flush=True)
if "rdf" in want:
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rdf_small.npz"))
    frames = z["frames"].astype(np.float64)
    t = _lib.DeviceTraj(frames.shape[1], frames.shape[0])
    t.upload(frames)
    out, read = t.rdf_cdf_sum(z["molid"], z["chain"], z["flags"], z["refidx"], com=False, step=float(z["step"]), end=float(z["end"]), cpus=1)
    assert read == frames.shape[0] and np.array_equal(out, z["cdf_frames_com0"].sum(axis=0))
    t.close()
    print("ok rdf", flush=True)
print("all ok")
