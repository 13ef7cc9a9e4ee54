"""Edge cases of chem.RotatorTranslatorToSuper pushed through EVERY kernel family on the device, against the oracle.

The reference repairs an improper rotation (det(U V^T) < 0, geometric.go:183-188) by flipping the axis of the smallest
singular value: R = U diag(1, 1, -1) V^T.  The device solver (csrc/kabsch3.cuh) reaches that branch through its one-sided
Jacobi fallback, taken whenever det H <= 0 or H is near-singular.  Three families of input exercise it:

  mirror   frames are mirror images of the template (det H < 0, all singular values well separated)
  planar   template and frames lie in a plane (H has rank 2: sigma_3 = 0); half the frames are mirrored inside the plane
  neardeg  mirror images of an anisotropic cloud whose two smallest singular values differ by a relative gap of 1e-3

and each goes through: the generic kernel, the cluster kernel with several cluster / stage / depth / residency shapes, the
warp-per-frame kernels (small and chunked), the gather kernels (small, chunked and CTA-per-frame), the split forms
(fit + elementwise sweep, small and large frames), the per-atom-sum passes, and the per-pair epilogue of the all-pairs GEMM.
Tolerances are BASELINE.json's: RMSD / translations 1e-9 A, rotations 1e-10.
"""
import numpy as np
import pytest

import oracle
from gochem_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL_RMSD = 1e-9
TOL_ROT = 1e-10
TOL_XYZ = 1e-9

KINDS = ["mirror", "planar", "neardeg"]


def _rot(rng):
    q = rng.standard_normal(4)
    return synth.quat_to_rot(q / np.linalg.norm(q))


def make_edge_case(kind, natoms, nframes, seed=0, noise=0.3):
    """(template, frames) whose correlation matrices sit on the reflection / rank-deficient branch."""
    rng = np.random.default_rng(1000 * seed + natoms)
    if kind == "planar":
        ref = np.zeros((natoms, 3))
        ref[:, :2] = rng.standard_normal((natoms, 2)) * [14.0, 9.0]
    elif kind == "neardeg":
        ref = rng.standard_normal((natoms, 3)) * [15.0, 8.0, 8.0]
        # principal axes of the SAMPLE, then the two smallest second moments set to differ by a relative 1e-3 exactly
        c = ref - ref.mean(0)
        _, vecs = np.linalg.eigh(c.T @ c)
        c = c @ vecs[:, ::-1]
        s = np.sqrt((c * c).sum(0))
        ref = c * [1.0, 1.0, s[1] / s[2] * (1.0 - 1e-3)]
    else:
        ref = synth.make_reference(natoms)
    frames = np.empty((nframes, natoms, 3))
    for f in range(nframes):
        x = ref.copy()
        if kind == "planar":
            x[:, :2] += rng.standard_normal((natoms, 2)) * noise
            if f % 2:
                x[:, 1] = -x[:, 1]          # mirrored inside the plane: a proper rotation (a flip) still superposes it
        elif kind == "neardeg":
            x += rng.standard_normal((natoms, 3)) * 1e-3 * noise
            x[:, 0] = -x[:, 0]              # mirror image: det H < 0, the flipped axis is one of the two near-equal ones
        else:
            x += rng.standard_normal((natoms, 3)) * noise
            if f % 4 != 3:                  # three frames in four are mirror images, the fourth is a plain copy
                x[:, 2] = -x[:, 2]
        frames[f] = x @ _rot(rng) + rng.uniform(-50, 50, 3)
    return ref, frames


def check_against_oracle(out, frames, ref, idx=None, rot_tol=TOL_ROT):
    rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, idx, nthreads=4)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
    assert np.abs(out["rot"] - rot).max() < rot_tol
    assert np.abs(out["trans1"] - t1).max() < TOL_XYZ and np.abs(out["trans2"] - t2).max() < TOL_XYZ
    # proper rotations, every one of them
    dets = np.linalg.det(out["rot"])
    assert np.abs(dets - 1.0).max() < 1e-12
    return rm, rot


def test_the_cases_really_are_on_the_reflection_branch():
    """Guard for the generator: det H < 0 for the mirrored frames, sigma_3 = 0 for the planar ones, a 1e-3 gap for neardeg."""
    for kind in KINDS:
        ref, frames = make_edge_case(kind, 200, 8)
        b = ref - ref.mean(0)
        neg = 0
        for f in range(8):
            a = frames[f] - frames[f].mean(0)
            h = a.T @ b
            s = np.linalg.svd(h, compute_uv=False)
            if kind == "planar":
                assert s[2] < 1e-9 * s[0]
            else:
                neg += np.linalg.det(h) < 0
            if kind == "neardeg":
                assert abs(s[1] - s[2]) / s[1] < 3e-3
        if kind == "mirror":
            assert neg == 6
        if kind == "neardeg":
            assert neg == 8


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("natoms", [40, 300, 1000, 2500])
@pytest.mark.parametrize("variant", [_lib.KERNEL_GENERIC, _lib.KERNEL_AUTO], ids=["generic", "auto"])
def test_fit_write_back_and_sums(kind, natoms, variant):
    """Generic kernel / the kernel KERNEL_AUTO picks for the size (warp-per-frame below 512 atoms, chunked to 1664, cluster above;
    split forms for write-back and per-atom sums on small structures): fit, subset fit, write-back, align.RMSDTraj sums."""
    nframes = 37
    ref, frames = make_edge_case(kind, natoms, nframes)
    t = _lib.DeviceTraj(natoms, nframes)
    t.set_kernel_variant(variant)
    t.upload(frames)
    try:
        check_against_oracle(t.super_rmsd(ref), frames, ref)
        idx = np.arange(1, natoms, 3)
        check_against_oracle(t.super_rmsd(ref, idx, idx), frames, ref, idx)
        # explicit residual for every frame (the reference's own way, geometric.go:284-286)
        t.set_rmsd_mode(1)
        check_against_oracle(t.super_rmsd(ref), frames, ref)
        t.set_rmsd_mode(0)
        # align.RMSDTraj's pass with write-back: per-atom sums and moved frames
        dev = t.peratom_dev_sum(ref, idx, write_back=True)
        want, total, aligned = oracle.rmsd_traj(frames, ref, idx, cpus=1, want_aligned=True)
        assert total == nframes and np.abs(dev / nframes - want).max() < 1e-10
        assert np.abs(t.download() - aligned).max() < TOL_XYZ
        # chem.Super semantics on all atoms
        t.upload(frames)
        out = t.super_rmsd(ref, write_back=True)
        work = frames.copy()
        rm, *_ = oracle.super_rmsd_traj(work, ref, None, write_back=True, nthreads=4)
        assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
        assert np.abs(t.download() - work).max() < TOL_XYZ
        # per-atom sums without write-back (the fused form and, on small structures, the split form)
        t.upload(frames)
        for thr in (-1, 0):
            t.set_split_thresholds(thr, thr)
            dev = t.peratom_dev_sum(ref, None)
            want, total = oracle.rmsd_traj(frames, ref, np.arange(natoms), cpus=1)
            assert np.abs(dev / nframes - want).max() < 1e-10
        t.set_split_thresholds(-1, -1)
    finally:
        t.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("natoms,cluster,stages,depth,resident", [
    (2500, 1, 0, 0, -1), (2500, 2, 2, 1, 0), (2500, 4, 3, 1, 1), (5000, 2, 0, 0, -1), (5000, 8, 4, 2, 1),
    (10000, 3, 2, 3, 0), (10000, 4, 3, 1, 1), (1500, 1, 4, 6, 0), (20000, 6, 2, 2, 0)])
def test_cluster_kernel_every_shape(kind, natoms, cluster, stages, depth, resident):
    nframes = 29
    ref, frames = make_edge_case(kind, natoms, nframes, seed=cluster)
    t = _lib.DeviceTraj(natoms, nframes)
    t.set_kernel_variant(_lib.KERNEL_CLUSTER)
    t.set_cluster_tuning(cluster, stages, depth, resident)
    t.upload(frames)
    try:
        check_against_oracle(t.super_rmsd(ref), frames, ref)
        plan = t.last_cluster_plan()
        assert plan["cluster"] == cluster
        idx = np.arange(0, natoms, 2)
        check_against_oracle(t.super_rmsd(ref, idx, idx), frames, ref, idx)
        dev = t.peratom_dev_sum(ref, idx, write_back=True)
        want, total, aligned = oracle.rmsd_traj(frames, ref, idx, cpus=1, want_aligned=True)
        assert np.abs(dev / nframes - want).max() < 1e-10
        assert np.abs(t.download() - aligned).max() < TOL_XYZ
    finally:
        t.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("natoms,nidx", [(9000, 300), (9000, 2000), (33000, 2500), (33000, 300)])
def test_gather_kernels(kind, natoms, nidx):
    """Subset fits that read only the subset's atoms: warp-per-frame gather (<= 512), its chunked form (<= 4096), and the
    CTA-per-frame gather kernel; then lists that pair DIFFERENT atoms and repeat a test atom (super_gather_kernel whatever the size)."""
    nframes = 23
    ref, frames = make_edge_case(kind, natoms, nframes)
    rng = np.random.default_rng(nidx)
    idx = np.sort(rng.choice(natoms, nidx, replace=False))
    # the SUBSET must sit on the branch too: rebuild the frames so that the subset of each frame is an edge case of its own
    sref, sframes = make_edge_case(kind, nidx, nframes, seed=7)
    ref[idx] = sref
    frames[:, idx, :] = sframes
    t = _lib.DeviceTraj(natoms, nframes)
    t.upload(frames)
    try:
        check_against_oracle(t.super_rmsd(ref, idx, idx), frames, ref, idx)
        t.set_kernel_variant(_lib.KERNEL_GENERIC)       # super_gather_kernel for every list shape
        check_against_oracle(t.super_rmsd(ref, idx, idx), frames, ref, idx)
        t.set_kernel_variant(_lib.KERNEL_AUTO)
        # two different lists with a repeated test atom: no plane layout exists, the gather kernel pairs them
        it = idx.copy()
        ir = idx.copy()
        it[1] = it[0]
        got = t.super_rmsd(ref, it, ir)
        for f in (0, 1, nframes - 1):
            _, r, t1, t2 = oracle.super_(frames[f], ref, it, ir, details=True)
            assert np.abs(got["rot"][f] - r).max() < TOL_ROT
            assert np.abs(got["trans1"][f] - t1).max() < TOL_XYZ and np.abs(got["trans2"][f] - t2).max() < TOL_XYZ
    finally:
        t.close()


@pytest.mark.parametrize("kind", KINDS)
def test_large_frame_split_path(kind):
    """Frames beyond the cluster kernel's resident capacity: fit first, then the elementwise sweeps (transform.cu)."""
    natoms, nframes, nidx = 24000, 9, 400
    ref, frames = make_edge_case(kind, natoms, nframes)
    idx = np.sort(np.random.default_rng(3).choice(natoms, nidx, replace=False))
    sref, sframes = make_edge_case(kind, nidx, nframes, seed=5)
    ref[idx] = sref
    frames[:, idx, :] = sframes
    t = _lib.DeviceTraj(natoms, nframes)
    t.upload(frames)
    try:
        dev = t.peratom_dev_sum(ref, idx)
        want, total = oracle.rmsd_traj(frames, ref, idx, cpus=1)
        assert np.abs(dev / nframes - want).max() < 1e-10
        cand = np.sort(np.random.default_rng(4).choice(natoms, 1000, replace=False))
        some = t.peratom_dev_sum_atoms(ref, idx, cand)
        assert np.abs(some / nframes - want[cand]).max() < 1e-10
        out = t.super_rmsd(ref, idx, idx, write_back=True)
        work = frames.copy()
        rm, rot, t1, t2 = oracle.super_rmsd_traj(work, ref, idx, write_back=True, nthreads=4)
        assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
        assert np.abs(t.download() - work).max() < TOL_XYZ
    finally:
        t.close()


@pytest.mark.parametrize("kind", KINDS)
def test_allpairs_fit_epilogue(kind):
    """The per-pair solve in the epilogue of the all-pairs GEMM (allpairs_fit_dmma_kernel): pairs of mirror images, planar frames,
    near-degenerate clouds."""
    natoms, nframes = 129, 70
    ref, frames = make_edge_case(kind, natoms, nframes)
    if kind == "mirror":
        frames[0::2] = make_edge_case("mirror", natoms, nframes, seed=1, noise=0.2)[1][0::2]
    t = _lib.DeviceTraj(natoms, nframes)
    t.upload(frames)
    try:
        got, _, _ = t.allpairs_rmsd(fit=True)
        assert np.array_equal(got, got.T) and (np.diag(got) == 0).all()
        rng = np.random.default_rng(0)
        nneg = 0
        for _ in range(150):
            i, j = rng.integers(0, nframes, 2)
            if i == j:
                continue
            a = frames[i] - frames[i].mean(0)
            b = frames[j] - frames[j].mean(0)
            nneg += np.linalg.det(a.T @ b) < 0
            s = oracle.super_(frames[i], frames[j])
            assert abs(got[i, j] - oracle.rmsd(s, frames[j])) < TOL_RMSD
        if kind == "mirror":
            assert nneg > 20        # the sample really holds reflected pairs
    finally:
        t.close()


def test_host_api_rotator_on_a_mirror_image():
    """chem.RotatorTranslatorToSuper / chem.Super of the host mirror (a batch of one on the device) on the reflection branch."""
    from gochem_b200 import host_api as H

    for kind in KINDS:
        ref, frames = make_edge_case(kind, 77, 3)
        for f in range(3):
            tr, r, a, b = H.RotatorTranslatorToSuper(frames[f], ref)
            otr, orr, oa, ob = oracle.rotator_translator_to_super(frames[f], ref)
            assert np.abs(r - orr).max() < TOL_ROT and abs(np.linalg.det(r) - 1) < 1e-12
            assert np.abs(tr - otr).max() < TOL_XYZ and np.abs(a - oa).max() < TOL_XYZ and np.abs(b - ob).max() < TOL_XYZ
            assert np.abs(H.Super(frames[f], ref) - oracle.super_(frames[f], ref)).max() < TOL_XYZ
