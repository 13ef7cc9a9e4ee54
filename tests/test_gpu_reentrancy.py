"""Reentrancy of the C ABI and a stress test of the cluster kernel's mbarrier / DSMEM protocol.

The reference's functions are reentrant on distinct matrices: align.RMSDTraj calls chem.Super from o.cpus goroutines at once
(align/lovo.go:290-295, 329).  Here every setting and every scratch buffer belongs to a trajectory handle, so calls on
DISTINCT handles may run concurrently from different threads (ctypes releases the GIL inside a call).

compute-sanitizer is refused by the pool (profiles/r1_s5_sanitize_target_plain.txt), so the stand-in for racecheck is the
second half: many launches of every cluster shape on the same input must give bit-identical results every time -- a lost or
early exchange (a partial sum read before it landed, a stage released too soon) shows up as a changed bit.
"""
import threading

import numpy as np
import pytest

import oracle
from gochem_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_two_handles_two_threads_different_settings():
    """Two handles on one device, each driven by its own thread with its own kernel variant, cluster shape, RMSD mode and split
    thresholds, and frame counts that make the per-handle scratch grow while the other handle's kernels are in flight."""
    cases = [
        dict(natoms=2500, nframes=300, variant=_lib.KERNEL_CLUSTER, tuning=(2, 3, 1, 1), explicit=1, seed=1),
        dict(natoms=700, nframes=420, variant=_lib.KERNEL_AUTO, tuning=(0, 0, 0, -1), explicit=0, seed=2),
        dict(natoms=5000, nframes=150, variant=_lib.KERNEL_GENERIC, tuning=(0, 0, 0, -1), explicit=0, seed=3),
        dict(natoms=3100, nframes=260, variant=_lib.KERNEL_CLUSTER, tuning=(4, 0, 2, 0), explicit=0, seed=4),
    ]
    work = []
    for c in cases:
        ref, frames = synth.make_case(c["natoms"], 16)
        frames = np.tile(frames, (c["nframes"] // 16 + 1, 1, 1))[: c["nframes"]]
        frames = frames + np.random.default_rng(c["seed"]).standard_normal(frames.shape) * 0.02
        idx = np.arange(0, c["natoms"], 3)
        rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, None, nthreads=8)
        rm_s, rot_s, _, _ = oracle.super_rmsd_traj(frames, ref, idx, nthreads=8)
        dev, total = oracle.rmsd_traj(frames, ref, idx, cpus=1, nthreads=8)
        work.append((c, ref, frames, idx, (rm, rot, t1, t2), (rm_s, rot_s), dev))
    errors = []
    start = threading.Barrier(len(work))

    def run(c, ref, frames, idx, full, sub, dev):
        try:
            t = _lib.DeviceTraj(c["natoms"], c["nframes"])
            t.set_kernel_variant(c["variant"])
            t.set_cluster_tuning(*c["tuning"])
            t.set_rmsd_mode(c["explicit"])
            start.wait()
            first = None
            for it in range(12):
                n = c["nframes"] if it % 3 else c["nframes"] // 2     # the squared-error scratch shrinks and grows
                t.upload(frames[:n])
                t.set_nframes(n)
                out = t.super_rmsd(ref)
                assert np.abs(out["rmsd"] - full[0][:n]).max() < 1e-9, "rmsd"
                assert np.abs(out["rot"] - full[1][:n]).max() < 1e-10, "rot"
                assert np.abs(out["trans1"] - full[2][:n]).max() < 1e-9 and np.abs(out["trans2"] - full[3][:n]).max() < 1e-9
                s = t.super_rmsd(ref, idx, idx)
                assert np.abs(s["rmsd"] - sub[0][:n]).max() < 1e-9 and np.abs(s["rot"] - sub[1][:n]).max() < 1e-10
                if n == c["nframes"]:
                    d = t.peratom_dev_sum(ref, idx)
                    assert np.abs(d / n - dev).max() < 1e-10, "per-atom sums"
                    if first is None:
                        first = (out["rmsd"].copy(), out["rot"].copy(), d.copy())
                    else:   # same handle, same input: the same bits, whatever the other threads were doing
                        assert np.array_equal(first[0], out["rmsd"]) and np.array_equal(first[1], out["rot"])
                        assert np.array_equal(first[2], d)
                plan = t.last_cluster_plan()
                if c["variant"] == _lib.KERNEL_CLUSTER and c["tuning"][0]:
                    assert plan["cluster"] == c["tuning"][0], "another handle's tuning leaked into this one"
            t.close()
        except Exception as exc:   # noqa: BLE001 - reported to the main thread
            errors.append((c, repr(exc)))
            try:
                start.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=run, args=w) for w in work]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_defaults_do_not_touch_existing_handles():
    """gcb_set_* change the defaults of handles created afterwards only."""
    lib = _lib.load()
    ref, frames = synth.make_case(2000, 20)
    a = _lib.DeviceTraj(2000, 20)
    a.upload(frames)
    try:
        _lib.check(lib.gcb_set_kernel_variant(_lib.KERNEL_CLUSTER))
        _lib.check(lib.gcb_set_cluster_tuning(4, 0, 0, -1))
        b = _lib.DeviceTraj(2000, 20)
        b.upload(frames)
        ra = a.super_rmsd(ref)          # AUTO: the chunked warp-per-frame kernel would not even make a cluster plan
        rb = b.super_rmsd(ref)
        assert b.last_cluster_plan()["cluster"] == 4
        assert a.last_cluster_plan()["cluster"] in (0, 1)
        assert np.abs(ra["rmsd"] - rb["rmsd"]).max() < 1e-9
        b.close()
    finally:
        lib.gcb_set_cluster_tuning(0, 0, 0, -1)
        lib.gcb_set_kernel_variant(_lib.KERNEL_AUTO)
        a.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cluster_protocol_stress_bit_identical(seed):
    """Random cluster / stage / depth / residency shapes x hundreds of launches each: rmsd, rotations, per-atom sums and the
    written-back frames must be the same bits on every launch (the stand-in for a racecheck run)."""
    rng = np.random.default_rng(100 + seed)
    launches = 0
    for _ in range(7):
        cluster = int(rng.integers(1, 9))
        resident = int(rng.integers(0, 2))
        natoms = int(rng.integers(1300, min(2400 * cluster, 19000)))
        stages = int(rng.choice([0, 2, 3, 4]))
        depth = int(rng.integers(1, 4 if resident else 7))
        nframes = int(rng.integers(700, 1600))
        ref, base = synth.make_case(natoms, 8)
        frames = np.tile(base, (nframes // 8 + 1, 1, 1))[:nframes]
        frames[1::7] = frames[0] + 1e-7                     # near-template frames: the explicit-residual path interleaved
        t = _lib.DeviceTraj(natoms, nframes)
        t.set_kernel_variant(_lib.KERNEL_CLUSTER)
        try:
            t.set_cluster_tuning(cluster, stages, depth, resident)
            t.upload(frames)
            idx = np.arange(0, natoms, 2)
            try:
                want = t.super_rmsd(ref)
            except _lib.GcbError:
                continue                                   # no plan for this random shape (too many atoms per CTA)
            want_dev = t.peratom_dev_sum(ref, idx)
            want_sub = t.super_rmsd(ref, idx, idx)
            reps = 60
            for _ in range(reps):
                got = t.super_rmsd(ref)
                assert np.array_equal(got["rmsd"], want["rmsd"]) and np.array_equal(got["rot"], want["rot"])
                assert np.array_equal(got["trans1"], want["trans1"])
                assert np.array_equal(t.peratom_dev_sum(ref, idx), want_dev)
                s = t.super_rmsd(ref, idx, idx)
                assert np.array_equal(s["rmsd"], want_sub["rmsd"]) and np.array_equal(s["rot"], want_sub["rot"])
            launches += 3 * reps
            # write-back twice from the same input gives the same frames
            t.super_rmsd(ref, write_back=True, want=())
            moved = t.download()
            t.upload(frames)
            t.super_rmsd(ref, write_back=True, want=())
            assert np.array_equal(t.download(), moved)
        finally:
            t.close()
    assert launches >= 360
