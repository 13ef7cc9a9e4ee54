"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances are BASELINE.json's: RMSD and translations within 1e-9 A, rotations within 1e-10
(entries are <= 1, so absolute = relative to the matrix norm), LOVO index sets identical.
"""
import os

import numpy as np
import pytest

import oracle
from gochem_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL_RMSD = 1e-9
TOL_ROT = 1e-10
TOL_XYZ = 1e-9

VARIANTS = [_lib.KERNEL_GENERIC, _lib.KERNEL_AUTO]


@pytest.fixture(params=VARIANTS, ids=["generic", "auto"])
def variant(request):
    _lib.check(_lib.load().gcb_set_kernel_variant(request.param))
    yield request.param
    _lib.load().gcb_set_kernel_variant(_lib.KERNEL_AUTO)


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def device_traj(frames):
    t = _lib.DeviceTraj(frames.shape[1], frames.shape[0])
    t.upload(frames)
    return t


@pytest.mark.parametrize("name", ["super_all_n50", "super_subset_n203", "super_cfg1_sample_n1500", "super_edge_mirror_n120", "super_edge_planar_n120",
                                  "super_edge_neardeg_n120"])
def test_golden_vectors(golden_dir, name, variant):
    g = load(golden_dir, name)
    ref = g["ref"].astype(np.float64)
    frames = g["frames"].astype(np.float64)
    idx = g["idx"] if g["idx"].size else None
    t = device_traj(frames)
    out = t.super_rmsd(ref, idx, idx)
    assert np.abs(out["rmsd"] - g["rmsd"]).max() < TOL_RMSD
    assert np.abs(out["rot"] - g["rot"]).max() < TOL_ROT
    assert np.abs(out["trans1"] - g["trans1"]).max() < TOL_XYZ
    assert np.abs(out["trans2"] - g["trans2"]).max() < TOL_XYZ
    # frames untouched without write_back
    assert np.array_equal(t.download(), frames)
    out2 = t.super_rmsd(ref, idx, idx, write_back=True)
    moved = t.download()
    assert np.abs(moved[1] - g["moved_first"]).max() < TOL_XYZ
    assert np.abs(moved[-1] - g["moved_last"]).max() < TOL_XYZ
    assert np.abs(out2["rmsd"] - g["rmsd"]).max() < TOL_RMSD
    t.close()


def test_config1_full_vs_oracle(variant):
    """BASELINE.json configs[0]: 1000 frames x 1500 atoms against frame 0."""
    ref, frames = synth.make_case(1500, 1000, frame0_is_ref=True)
    t = device_traj(frames)
    out = t.super_rmsd(ref, write_back=True)
    moved = t.download()
    rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, None, write_back=True, nthreads=8)  # frames now superposed
    assert out["rmsd"][0] < 1e-12  # frame 0 is the reference: explicit residual, no cancellation
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
    assert np.abs(out["rot"] - rot).max() < TOL_ROT
    assert np.abs(out["trans1"] - t1).max() < TOL_XYZ
    assert np.abs(out["trans2"] - t2).max() < TOL_XYZ
    assert np.abs(moved - frames).max() < TOL_XYZ
    t.close()


@pytest.mark.parametrize("natoms", [3, 17, 255, 256, 257, 1000, 4097])
def test_ragged_sizes(natoms, variant):
    ref, frames = synth.make_case(natoms, 5)
    t = device_traj(frames)
    out = t.super_rmsd(ref)
    rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, None)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
    assert np.abs(out["rot"] - rot).max() < (TOL_ROT if natoms > 3 else 1e-8)
    t.close()


def test_weighted_subset_with_duplicates(variant):
    ref, frames = synth.make_case(300, 7)
    idx = np.array(list(range(0, 300, 3)) + [3, 3, 9])  # repeated atoms count twice (SomeVecs copies rows)
    t = device_traj(frames)
    out = t.super_rmsd(ref, idx, idx, write_back=True)
    moved = t.download()
    for f in range(7):
        s, r, a, b = oracle.super_(frames[f], ref, idx, idx, details=True)
        assert np.abs(moved[f] - s).max() < TOL_XYZ
        assert np.abs(out["rot"][f] - r).max() < TOL_ROT
        assert abs(out["rmsd"][f] - oracle.rmsd(s, ref, idx, idx)) < TOL_RMSD
    t.close()


def test_gather_lists_and_smaller_template(variant):
    ref, frames = synth.make_case(120, 4)
    rng = np.random.default_rng(0)
    it = rng.permutation(120)[:40]
    ir = it.copy()
    t = device_traj(frames)
    # same pairs expressed through a permuted, smaller template: templa = ref[ir], lists (it, 0..39)
    templa = ref[ir]
    out = t.super_rmsd(templa, it, np.arange(40), write_back=True)
    moved = t.download()
    for f in range(4):
        s, r, a, b = oracle.super_(frames[f], templa, it, np.arange(40), details=True)
        assert np.abs(moved[f] - s).max() < TOL_XYZ
        assert np.abs(out["rot"][f] - r).max() < TOL_ROT
        assert abs(out["rmsd"][f] - oracle.rmsd(s, templa, it, np.arange(40))) < TOL_RMSD
    # genuinely different pairing
    t.upload(frames)
    ir2 = rng.permutation(120)[:40]
    out = t.super_rmsd(ref, it, ir2)
    for f in range(4):
        s, r, a, b = oracle.super_(frames[f], ref, it, ir2, details=True)
        assert np.abs(out["rot"][f] - r).max() < 1e-8  # ill-conditioned on purpose (random pairing)
        assert abs(out["rmsd"][f] - oracle.rmsd(s, ref, it, ir2)) < 1e-8
    # chem.RMSD without a fit through the same two lists
    nf = t.rmsd(ref, it, ir2)
    for f in range(4):
        assert abs(nf[f] - oracle.mem_rmsd(frames[f], ref, it, ir2)) < TOL_RMSD
    # a test atom used twice, paired with different template atoms: the pairing cannot be folded into the template planes
    it3 = np.array([5, 5, 9, 33, 64, 64, 100, 2, 17, 80, 41, 42])
    ir3 = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
    out = t.super_rmsd(ref, it3, ir3)
    nf = t.rmsd(ref, it3, ir3)
    for f in range(4):
        s, r, a, b = oracle.super_(frames[f], ref, it3, ir3, details=True)
        assert np.abs(out["rot"][f] - r).max() < 1e-8
        assert abs(out["rmsd"][f] - oracle.rmsd(s, ref, it3, ir3)) < 1e-8
        assert abs(nf[f] - oracle.mem_rmsd(frames[f], ref, it3, ir3)) < TOL_RMSD
    t.close()


def test_rmsd_without_fit(variant):
    ref, frames = synth.make_case(500, 6)
    idx = np.arange(10, 400, 7)
    t = device_traj(frames)
    got = t.rmsd(ref)
    got_idx = t.rmsd(ref, idx, idx)
    got_pairs = t.rmsd(ref, idx, idx[::-1].copy())
    for f in range(6):
        assert abs(got[f] - oracle.rmsd(frames[f], ref)) < 1e-9 * max(1, got[f])
        assert abs(got_idx[f] - oracle.rmsd(frames[f], ref, idx, idx)) < 1e-9 * max(1, got_idx[f])
        assert abs(got_pairs[f] - oracle.rmsd(frames[f], ref, idx, idx[::-1].copy())) < 1e-9 * max(1, got_pairs[f])
    t.close()


def test_host_layouts_and_scale():
    ref, frames = synth.make_case(77, 5, through_f32=True)
    f32 = frames.astype(np.float32)
    t = _lib.DeviceTraj(77, 5)
    t.upload(f32, layout=_lib.F32_AOS)
    assert np.array_equal(t.download(), frames)
    soa = np.ascontiguousarray(f32.transpose(0, 2, 1))  # DCD: x[], y[], z[] per frame
    t.upload(soa, layout=_lib.F32_SOA)
    assert np.array_equal(t.download(), frames)
    t.upload(f32, layout=_lib.F32_AOS, scale=10.0)  # xtc nm -> A, 10*float64(c) (xtc.go:131)
    assert np.array_equal(t.download(), 10.0 * frames)
    # partial window
    t.upload(frames[:2], first=3)
    d = t.download()
    assert np.array_equal(d[3:], frames[:2])
    # STF fixed point: int32 per coordinate, float64(f) / 100 on the device (traj/stf/stf.go:341), bit for bit
    ints = np.rint(frames * 100.0).astype(np.int32)
    ints[0, 0] = [2147483647, -2147483648, 7]
    t.upload(ints, layout=_lib.I32_AOS, scale=100.0)
    assert np.array_equal(t.download(), ints.astype(np.float64) / 100.0)
    with pytest.raises(_lib.GcbError):
        t.upload(ints, layout=_lib.I32_AOS, scale=0.0)
    t.close()


def test_pinned_async_upload():
    ref, frames = synth.make_case(64, 8, through_f32=True)
    t = _lib.DeviceTraj(64, 8)
    pb = [_lib.PinnedBuffer(4 * 64 * 3 * 4) for _ in range(2)]
    f32 = frames.astype(np.float32)
    for k in range(2):
        v = pb[k].view(np.float32, (4, 64, 3))
        v[...] = f32[4 * k:4 * k + 4]
        t.upload_async(v, 4, 4 * k, _lib.F32_AOS, 1.0, k)
    out = t.super_rmsd(ref)
    t.wait(0); t.wait(1)
    rm, *_ = oracle.super_rmsd_traj(frames, ref, None)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
    for p in pb:
        p.free()
    t.close()


@pytest.mark.parametrize("natoms", [3, 33, 100, 257, 440, 512, 513, 1500, 2048])
def test_small_structure_kernel(natoms):
    """Structures of at most 512 atoms take the warp-per-frame kernel under KERNEL_AUTO, up to 2048 atoms its chunked form (the
    per-atom sums of those go to the cluster kernel): every mode against the oracle, with a frame count that is not a multiple of
    the 32-frame batches and a near-template frame that needs the explicit residual."""
    nframes = 75
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    frames[40] = frames[0] + 1e-6                      # almost the template: the one-pass RMSD is not allowed here
    t = device_traj(frames)
    out = t.super_rmsd(ref)
    rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, None)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
    if natoms >= 33:
        assert np.abs(out["rot"] - rot).max() < TOL_ROT
    assert np.abs(out["trans1"] - t1).max() < TOL_XYZ and np.abs(out["trans2"] - t2).max() < TOL_XYZ
    if natoms >= 33:
        idx = np.arange(1, natoms, 3)
        sub = t.super_rmsd(ref, idx, idx)
        rm_s, rot_s, _, _ = oracle.super_rmsd_traj(frames, ref, idx)
        assert np.abs(sub["rmsd"] - rm_s).max() < TOL_RMSD and np.abs(sub["rot"] - rot_s).max() < TOL_ROT
        dev = t.peratom_dev_sum(ref, idx, write_back=True)
        want, total, aligned = oracle.rmsd_traj(frames, ref, idx, cpus=5, want_aligned=True)
        assert total == nframes and np.abs(dev / nframes - want).max() < 1e-10
        assert np.abs(t.download() - aligned).max() < TOL_XYZ
        t.upload(frames)
    t.set_rmsd_mode(1)                                 # explicit residual for every frame
    ex = t.super_rmsd(ref, write_back=True)
    t.set_rmsd_mode(0)
    assert np.abs(ex["rmsd"] - rm).max() < TOL_RMSD
    moved = t.download()
    for f in (0, 31, 32, 74):
        assert np.abs(moved[f] - oracle.super_(frames[f], ref)).max() < TOL_XYZ
    t.close()


def test_upload_of_raw_records():
    """gcb_traj_upload_records_async: float32 planes inside fixed-size records (raw DCD frames: markers, a unit-cell block, a
    trailing block), picked apart on the device; bad layouts are argument errors."""
    natoms, nframes = 77, 9
    ref, frames = synth.make_case(natoms, nframes, through_f32=True)
    plane = 4 * natoms
    off = (4 + 48 + 4 + 4, 0, 0)
    off = (off[0], off[0] + plane + 8, off[0] + 2 * (plane + 8))
    rec = off[2] + plane + 4 + (4 + plane + 4)
    pb = _lib.PinnedBuffer(rec * nframes)
    pb.u8[:] = 0xAB                                   # whatever sits between the planes must not matter
    f32 = frames.astype(np.float32)
    for f in range(nframes):
        for c in range(3):
            pb.u8[f * rec + off[c]: f * rec + off[c] + plane] = f32[f, :, c].copy().view(np.uint8)
    t = _lib.DeviceTraj(natoms, nframes + 3)
    lib = _lib.load()
    _lib.check(lib.gcb_traj_upload_records_async(t.h, 3, pb.ptr, nframes, rec, off[0], off[1], off[2], 1.0, 1))
    t.wait(1)
    assert lib.gcb_traj_stage_bytes(t.h) >= 64 << 20
    back = t.download(3, nframes)
    assert np.array_equal(back, frames)
    _lib.check(lib.gcb_traj_upload_records_async(t.h, 0, pb.ptr, 2, rec, off[0], off[1], off[2], 10.0, 0))   # the xtc-style scale
    t.wait(0)
    assert np.array_equal(t.download(0, 2), 10.0 * frames[:2])
    for bad in [(rec, off[0] + 2, off[1], off[2]), (rec, off[0], off[1], rec - 8), (0, 0, 0, 0), (rec + 2, off[0], off[1], off[2])]:
        assert lib.gcb_traj_upload_records_async(t.h, 0, pb.ptr, 1, bad[0], bad[1], bad[2], bad[3], 1.0, 0) == _lib.ERR_ARG
    pb.free()
    t.close()


def test_peratom_dev_sum_matches_rmsdtraj(variant):
    ref, frames = synth.make_case(333, 24)
    idx = np.arange(0, 333, 2)
    t = device_traj(frames)
    got = t.peratom_dev_sum(ref, idx)
    want, total = oracle.rmsd_traj(frames, ref, idx, cpus=4)
    assert total == 24
    assert np.abs(got / 24.0 - want).max() < 1e-10
    # all atoms, write back the aligned trajectory (writeTraj path, lovo.go:337-339)
    t.upload(frames)
    got = t.peratom_dev_sum(ref, None, write_back=True)
    want, total, aligned = oracle.rmsd_traj(frames, ref, np.arange(333), cpus=4, want_aligned=True)
    assert np.abs(got / 24.0 - want).max() < 1e-10
    assert np.abs(t.download() - aligned).max() < TOL_XYZ
    t.close()


def test_errors_mirror_the_reference():
    ref, frames = synth.make_case(50, 3)
    t = device_traj(frames)
    with pytest.raises(_lib.GcbError) as e:
        t.super_rmsd(ref[:40])  # "Ill formed coordinates for Superposition"
    assert e.value.code == _lib.ERR_SHAPE
    with pytest.raises(_lib.GcbError) as e:
        t.super_rmsd(ref, np.array([1, 2, 99]), np.array([1, 2, 3]))
    assert e.value.code == _lib.ERR_INDEXES
    with pytest.raises(_lib.GcbError) as e:
        t.super_rmsd(ref, first=2, nframes=5)
    assert e.value.code == _lib.ERR_ARG
    bad = frames.copy()
    bad[1, 7, 2] = np.nan
    t.upload(bad)
    with pytest.raises(_lib.GcbError) as e:
        t.super_rmsd(ref)  # mat.SVD failed
    assert e.value.code == _lib.ERR_SVD
    t.upload(frames)
    assert np.isfinite(t.super_rmsd(ref)["rmsd"]).all()  # the error flag does not stick
    t.close()


def test_empty_window_is_a_noop():
    ref, frames = synth.make_case(20, 2)
    t = device_traj(frames)
    out = t.super_rmsd(ref, first=1, nframes=0)
    assert out["rmsd"].size == 0
    t.close()


def test_device_synth_properties_full_size(variant):
    """Size-independent properties on a device-generated trajectory with the shape of
    BASELINE.json configs[1] scaled to fit the test budget (10k atoms; 20k frames = 4.8 GB)."""
    natoms, nframes = 10000, 20000
    ref = synth.make_reference(natoms)
    sig = synth.atom_sigmas(natoms)
    t = _lib.DeviceTraj(natoms, nframes)
    t.synth_fill(ref, sig, nframes, frame0_is_ref=True)
    out = t.super_rmsd(ref)
    assert out["rmsd"][0] < 1e-12
    assert (out["rmsd"][1:] > 0.5).all() and (out["rmsd"] < 5).all()
    # sampled frames against the oracle, on the exact bytes the GPU holds
    for f in [0, 1, 4999, 19999]:
        fr = t.download(f, 1)[0]
        s, r, a, b = oracle.super_(fr, ref, details=True)
        assert abs(out["rmsd"][f] - oracle.rmsd(s, ref)) < TOL_RMSD
        assert np.abs(out["rot"][f] - r).max() < TOL_ROT
        assert np.abs(out["trans1"][f] - a).max() < TOL_XYZ
    # rotations are proper and orthogonal
    rr = out["rot"]
    assert np.abs(np.einsum("fij,fkj->fik", rr, rr) - np.eye(3)).max() < 1e-13
    assert np.abs(np.linalg.det(rr) - 1).max() < 1e-13
    # Super in place, then: RMSD without fit equals the fitted RMSD; a second Super is the identity
    out_wb = t.super_rmsd(ref, write_back=True)
    assert np.abs(out_wb["rmsd"] - out["rmsd"]).max() < 1e-12
    nofit = t.rmsd(ref)
    assert np.abs(nofit - out["rmsd"]).max() < TOL_RMSD
    again = t.super_rmsd(ref)
    assert np.abs(again["rot"] - np.eye(3)).max() < 1e-10
    assert np.abs(again["rmsd"] - out["rmsd"]).max() < TOL_RMSD
    # rigid copies of the reference give zero
    t.synth_fill(ref, np.zeros(natoms), 2000)
    z = t.super_rmsd(ref, nframes=2000)
    assert z["rmsd"].max() < 1e-11
    t.close()


@pytest.mark.parametrize("natoms,cluster,stages,depth,resident", [
    (1500, 1, 0, 0, 1), (1500, 1, 3, 1, 1), (1500, 2, 0, 3, 1), (1500, 4, 4, 2, 1), (1500, 8, 0, 1, 1),
    (5000, 2, 0, 0, 1), (5000, 4, 0, 2, 1), (5000, 8, 5, 3, 1), (10000, 4, 0, 0, 1), (10000, 8, 0, 3, 1),
    (613, 1, 0, 0, 1), (613, 8, 0, 2, 1), (40, 4, 0, 1, 1),
    (1500, 1, 0, 0, 0), (1500, 1, 2, 6, 0), (1500, 3, 0, 4, 0), (5000, 2, 0, 5, 0), (5000, 5, 3, 2, 0),
    (10000, 4, 0, 0, 0), (10000, 6, 0, 6, 0), (10000, 7, 2, 1, 0), (10000, 8, 0, 4, 0), (613, 8, 0, 3, 0), (40, 4, 0, 1, 0),
])
def test_cluster_kernel_tunings(natoms, cluster, stages, depth, resident):
    """Every cluster shape / pipeline depth must give the same answers as the oracle."""
    lib = _lib.load()
    nframes = 333
    ref, frames = synth.make_case(natoms, 16)
    frames = np.tile(frames, (21, 1, 1))[:nframes]
    frames += np.random.default_rng(1).standard_normal(frames.shape) * 0.01
    idx = np.arange(0, natoms, 3)
    t = device_traj(frames)
    try:
        t.set_kernel_variant(_lib.KERNEL_CLUSTER)
        t.set_cluster_tuning(cluster, stages, depth, resident)
        out = t.super_rmsd(ref)
        rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, None, nthreads=8)
        assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD
        assert np.abs(out["rot"] - rot).max() < TOL_ROT
        assert np.abs(out["trans1"] - t1).max() < TOL_XYZ
        # weighted subset + LOVO sums + write-back in one pass
        dev = t.peratom_dev_sum(ref, idx, write_back=True)
        want, total, aligned = oracle.rmsd_traj(frames, ref, idx, cpus=3, want_aligned=True, nthreads=3)
        assert total == nframes
        assert np.abs(dev / nframes - want).max() < 1e-10
        assert np.abs(t.download() - aligned).max() < TOL_XYZ
    finally:
        t.close()


def _pairwise(frames, idx=None):
    x = frames if idx is None else frames[:, idx]
    d = x[:, None, :, :] - x[None, :, :, :]
    return np.sqrt((d * d).sum(axis=(2, 3)) / x.shape[1])


@pytest.mark.parametrize("natoms,nframes", [(50, 37), (333, 300), (1000, 129)])
def test_allpairs_rmsd_matrix(natoms, nframes):
    """BASELINE configs[3] at sizes numpy can check: D_ij = chem.RMSD(X_i, X_j), frames superposed first."""
    ref, frames = synth.make_case(natoms, nframes)
    frames[5] = frames[2]                      # exact duplicate
    frames[7] = frames[3] + 1e-7               # near duplicate: the expansion alone would lose it
    t = device_traj(frames)
    t.super_rmsd(ref, write_back=True)
    aligned = t.download()
    got, ms, redone = t.allpairs_rmsd()
    want = _pairwise(aligned)
    assert got.shape == (nframes, nframes)
    assert np.abs(got - want).max() < 1e-9
    assert np.array_equal(got, got.T) and (np.diag(got) == 0).all()
    assert got[2, 5] < 1e-12 and redone >= 2
    # D_ij against the single-pair entry point
    assert abs(got[1, 4] - oracle.rmsd(aligned[1], aligned[4])) < 1e-9
    # atom subset and a row block (what one rank of a multi-GPU run computes)
    idx = np.arange(0, natoms, 3)
    sub, _, _ = t.allpairs_rmsd(idx=idx, row0=10, nrows=20)
    assert np.abs(sub - _pairwise(aligned, idx)[10:30]).max() < 1e-9
    t.close()


def test_allpairs_last_wave_as_half_blocks_is_bit_identical():
    """17 block rows = 153 blocks: on 148 SMs the last wave (5 blocks) is computed as half blocks of 64 x 128
    (allpairs_dmma_half_kernel).  Same bits as without the split, and the rows agree with numpy; the last block row holds 52
    frames only, so the lower half of its diagonal block lies outside the matrix."""
    natoms, nframes = 40, 2100
    ref, frames = synth.make_case(natoms, nframes)
    frames[2099] = frames[11] + 1e-7
    t = device_traj(frames)
    t.super_rmsd(ref, write_back=True)
    aligned = t.download()
    got, _, redone = t.allpairs_rmsd()
    os.environ["GCB_AP_NO_SPLIT"] = "1"
    try:
        plain, _, _ = t.allpairs_rmsd()
    finally:
        del os.environ["GCB_AP_NO_SPLIT"]
    assert np.array_equal(got, plain)
    assert np.array_equal(got, got.T) and (np.diag(got) == 0).all() and redone >= 1
    for i in (0, 63, 64, 127, 1990, 2047, 2048, 2099):
        d = aligned - aligned[i][None]
        want = np.sqrt((d * d).sum(axis=(1, 2)) / natoms)
        assert np.abs(got[i] - want).max() < 1e-9
    t.close()


@pytest.mark.parametrize("fit", [False, True])
def test_allpairs_sharded_entry_single_rank(fit):
    """The multi-GPU entry point with a one-rank communicator: same tiles, same owner-directed epilogue, one owner."""
    natoms, nframes = 211, 301
    ref, frames = synth.make_case(natoms, nframes)
    frames[9] = frames[4] + 1e-7
    t = device_traj(frames)
    t.super_rmsd(ref, write_back=True)
    comm = _lib.Comm(0, 1, 0, _lib.Comm.unique_id())
    got, row0, gms, tms, redone = t.allpairs_rmsd_sharded(comm, fit=fit)
    want, _, _ = t.allpairs_rmsd(fit=fit)
    assert row0 == 0 and got.shape == (nframes, nframes)
    assert np.array_equal(got, want)
    if not fit:
        assert np.abs(got - _pairwise(t.download())).max() < 1e-9
    # a second call reuses the communicator's block
    again, _, _, _, _ = t.allpairs_rmsd_sharded(comm, fit=fit)
    assert np.array_equal(again, got)
    comm.close()
    t.close()


@pytest.mark.parametrize("natoms", [1500, 10000])
def test_one_pass_rmsd_agrees_with_explicit_residual(natoms):
    """Fit-only calls take the RMSD from the inner products of the single sweep when that is provably accurate and
    from the explicit residual otherwise; across noise levels from 1e-7 A to 3 A both must agree with the oracle."""
    lib = _lib.load()
    ref = synth.make_reference(natoms)
    rng = np.random.default_rng(5)
    scales = np.concatenate([[0.0], np.logspace(-7, 0.5, 63)])
    frames = np.empty((scales.size, natoms, 3))
    for k, s in enumerate(scales):
        q = rng.standard_normal(4); q /= np.linalg.norm(q)
        frames[k] = (ref + rng.standard_normal((natoms, 3)) * s) @ synth.quat_to_rot(q) + rng.uniform(-50, 50, 3)
    want, *_ = oracle.super_rmsd_traj(frames, ref, None, nthreads=8)
    t = device_traj(frames)
    try:
        t.set_rmsd_mode(0)
        one = t.super_rmsd(ref)["rmsd"]
        t.set_rmsd_mode(1)
        two = t.super_rmsd(ref)["rmsd"]
    finally:
        t.close()
    assert np.abs(two - want).max() < 1e-11          # explicit residual: rounding only
    assert np.abs(one - want).max() < 5e-10          # one-pass where its bound allows, explicit elsewhere
    assert one[0] < 1e-12                            # the template against itself goes through the explicit path
    big = want > 1.0
    assert big.sum() > 3 and np.abs(one[big] - want[big]).max() < 1e-10
    t.close()


@pytest.mark.parametrize("natoms", [1, 2, 3, 4])
def test_degenerate_atom_counts_do_not_nan(natoms):
    """Fewer than 3 atoms make the rotation non-unique (SURVEY.md §7 'degenerate inputs'); the solver must still
    return a proper rotation and the RMSD of the best fit."""
    rng = np.random.default_rng(natoms)
    ref = rng.standard_normal((natoms, 3)) * 5
    frames = rng.standard_normal((6, natoms, 3)) * 5
    t = device_traj(frames)
    out = t.super_rmsd(ref, write_back=True)
    assert np.isfinite(out["rmsd"]).all() and np.isfinite(out["rot"]).all()
    rr = out["rot"]
    assert np.abs(np.einsum("fij,fkj->fik", rr, rr) - np.eye(3)).max() < 1e-12
    assert (np.linalg.det(rr) > 0.999).all()
    want, *_ = oracle.super_rmsd_traj(frames, ref, None)
    assert np.abs(out["rmsd"] - want).max() < 1e-9      # the minimum is unique even when the rotation is not
    t.close()


def test_single_frame_and_larger_than_cluster_capacity():
    # one frame: the persistent kernel launches a single cluster
    ref, frames = synth.make_case(2000, 1)
    t = device_traj(frames)
    out = t.super_rmsd(ref)
    want, rot, *_ = oracle.super_rmsd_traj(frames, ref, None)
    assert abs(out["rmsd"][0] - want[0]) < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
    t.close()
    # 30,000 atoms exceed what 8 CTAs x 256 threads x 14 slots hold: the library falls back to the generic kernel
    ref, frames = synth.make_case(30000, 3)
    t = device_traj(frames)
    out = t.super_rmsd(ref, write_back=True)
    want, rot, *_ = oracle.super_rmsd_traj(frames, ref, None, nthreads=4)
    assert np.abs(out["rmsd"] - want).max() < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
    dev = t.peratom_dev_sum(ref)
    assert np.isfinite(dev).all()
    t.close()


@pytest.mark.parametrize("natoms,nframes", [(60, 45), (400, 100)])
def test_allpairs_rmsd_matrix_with_per_pair_fit(natoms, nframes):
    """Variant (ii) of BASELINE configs[3]: every pair optimally superposed first, D_ij = RMSD(Super(X_i, X_j), X_j)."""
    ref, frames = synth.make_case(natoms, nframes)
    frames[5] = frames[2] @ synth.quat_to_rot(np.array([0.5, 0.5, 0.5, 0.5])) + 3.0     # rigid copy: distance 0 after the fit
    t = device_traj(frames)
    got, ms, redone = t.allpairs_rmsd(fit=True)
    assert np.array_equal(got, got.T) and (np.diag(got) == 0).all()
    assert got[2, 5] < 1e-10 and redone >= 1
    rng = np.random.default_rng(0)
    pairs = [(1, 4), (0, nframes - 1), (7, 33)] + [tuple(rng.integers(0, nframes, 2)) for _ in range(60)]
    for i, j in pairs:
        if i == j:
            continue
        s = oracle.super_(frames[i], frames[j])
        assert abs(got[i, j] - oracle.rmsd(s, frames[j])) < 1e-9
    # fitted distances never exceed the unfitted ones of the same frames
    plain, _, _ = t.allpairs_rmsd()
    assert (got <= plain + 1e-9).all()
    # subset of atoms, row block
    idx = np.arange(0, natoms, 2)
    sub, _, _ = t.allpairs_rmsd(idx=idx, row0=3, nrows=9, fit=True)
    for i, j in [(3, 0), (5, 40), (11, 12)]:
        s = oracle.super_(frames[i][idx], frames[j][idx])
        assert abs(sub[i - 3, j] - oracle.rmsd(s, frames[j][idx])) < 1e-9
    t.close()


def test_center_of_mass_and_moment_tensor_batched():
    """chem.CenterOfMass / chem.MomentTensor for every resident frame."""
    natoms, nframes = 777, 40
    ref, frames = synth.make_case(natoms, nframes)
    rng = np.random.default_rng(1)
    mass = rng.uniform(1.0, 32.0, natoms)
    idx = np.arange(3, natoms, 5)
    t = device_traj(frames)
    for m, ix in [(None, None), (mass, None), (mass, idx), (None, idx)]:
        cen, mom = t.moments(m, ix)
        for f in (0, 1, nframes - 1):
            a = frames[f] if ix is None else frames[f][ix]
            mm = None if m is None else (m if ix is None else m[ix])
            oc, om = oracle.center_and_moment(a, mm)
            assert np.abs(cen[f] - oc).max() < 1e-10
            assert np.abs(mom[f] - om).max() < 1e-10 * np.abs(om).max()
    t.close()
    from gochem_b200 import host_api as H
    oc, om = oracle.center_and_moment(frames[3], mass)
    hc, hm = H.CenterAndMoment(frames[3], mass)
    assert np.abs(hc - oc).max() < 1e-10 and np.abs(hm - om).max() < 1e-10 * np.abs(om).max()
    oc, om = oracle.center_and_moment(frames[3])
    hc, hm = H.CenterAndMoment(frames[3])
    assert np.abs(hc - oc).max() < 1e-10 and np.abs(hm - om).max() < 1e-10 * np.abs(om).max()


def test_moment_tensor_of_a_small_selection_far_from_the_origin():
    """A ligand of 30 atoms (1 A across) at 400 A from the coordinate origin, in a frame of 6000 atoms: the reference centres the
    coordinates before it forms the tensor (geometric.go:719-731), so its result does not lose (|c| / extent)^2 * eps; the kernel
    takes its second moments about the first selected atom for the same reason."""
    rng = np.random.default_rng(4)
    natoms, nframes = 6000, 9
    frames = rng.normal(0, 30, (nframes, natoms, 3))
    lig = np.arange(5100, 5130)
    frames[:, lig, :] = np.array([400.0, -390.0, 410.0]) + rng.normal(0, 0.5, (nframes, lig.size, 3))
    mass = rng.uniform(1.0, 16.0, natoms)
    t = device_traj(frames)
    cen, mom = t.moments(mass, lig)
    t.close()
    for f in range(nframes):
        oc, om = oracle.center_and_moment(frames[f][lig], mass[lig])
        assert np.abs(cen[f] - oc).max() < 1e-10
        assert np.abs(mom[f] - om).max() < 1e-12 * np.abs(om).max()     # S2 - M c c^T from raw coordinates is good to 1e-10 here


@pytest.mark.parametrize("natoms,nframes", [(1, 3), (17, 5), (2048, 300), (2049, 7), (5003, 333), (10000, 150)])
def test_moments_stream_kernel_chunked_frames(natoms, nframes):
    """Frames of one chunk, several chunks with a ragged last chunk, more frames than CTAs: every frame checked."""
    ref, frames = synth.make_case(natoms, nframes)
    mass = np.random.default_rng(natoms).uniform(1.0, 32.0, natoms)
    t = device_traj(frames)
    cen, mom = t.moments(mass)
    t.close()
    msum = mass.sum()
    oc = (frames * mass[None, :, None]).sum(axis=1) / msum
    d = frames - oc[:, None, :]
    om = np.einsum("fi,fia,fib->fab", np.broadcast_to(mass, (nframes, natoms)), d, d)
    assert np.abs(cen - oc).max() < 1e-10
    # one sweep: M = sum m x x^T - (sum m) c c^T, so the rounding error scales with sum m |x|^2 (a single atom has M = 0)
    scale = (mass[None, :, None] * frames ** 2).sum(axis=(1, 2)).max()
    assert np.abs(mom - om).max() < 1e-10 * np.abs(om).max() + 1e-14 * scale
    c0, m0 = oracle.center_and_moment(frames[nframes - 1], mass)
    assert np.abs(cen[-1] - c0).max() < 1e-10 and np.abs(mom[-1] - m0).max() < 1e-10 * np.abs(m0).max() + 1e-14 * scale


@pytest.mark.parametrize("nidx", [1, 30, 257, 512, 700, 3000, 4096])
def test_small_subset_of_a_large_frame(nidx):
    """A fit on a subset of at most a quarter of the atoms that leaves the frames untouched reads only those atoms (gather): the
    warp-per-frame form up to 512 atoms, the generic gather kernel above; same results as the streaming kernels and the oracle."""
    natoms, nframes = (9000, 40) if nidx <= 512 else (33000, 37)   # sizes at which the gather forms are chosen (api.cu: launch_super)
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    idx = np.sort(np.random.default_rng(nidx).choice(natoms, nidx, replace=False))
    t = device_traj(frames)
    t.set_kernel_variant(_lib.KERNEL_AUTO)
    got = t.super_rmsd(ref, idx, idx)
    if nidx >= 30:
        rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, idx)
        assert np.abs(got["rmsd"] - rm).max() < TOL_RMSD and np.abs(got["rot"] - rot).max() < TOL_ROT
        assert np.abs(got["trans1"] - t1).max() < TOL_XYZ and np.abs(got["trans2"] - t2).max() < TOL_XYZ
    t.set_kernel_variant(_lib.KERNEL_CLUSTER if natoms <= 28672 else _lib.KERNEL_GENERIC)
    ref_path = t.super_rmsd(ref, idx, idx)           # the streaming kernel with the subset as a mask
    assert np.abs(got["rmsd"] - ref_path["rmsd"]).max() < TOL_RMSD
    if nidx >= 30:
        assert np.abs(got["rot"] - ref_path["rot"]).max() < TOL_ROT
    assert np.array_equal(t.download(), frames)       # untouched
    t.close()


def test_rmsd_without_fit_on_a_small_subset_of_a_large_frame():
    """chem.RMSD (no fit) over a small subset of a 20 000-atom frame takes the gather kernel; same numbers as the oracle."""
    natoms, nframes = 20000, 9
    ref, frames = synth.make_case(natoms, nframes)
    idx = np.sort(np.random.default_rng(2).choice(natoms, 400, replace=False))
    t = device_traj(frames)
    got = t.rmsd(ref, idx, idx)
    for f in range(nframes):
        assert abs(got[f] - oracle.rmsd(frames[f], ref, idx, idx)) < 1e-9 * max(1, got[f])
    t.close()


@pytest.mark.parametrize("natoms,nidx", [(33000, 0), (33000, 300), (33000, 5000), (40010, 977), (24000, 0), (24000, 450)])
def test_large_frames_write_back_and_per_atom_sums(natoms, nidx):
    """Frames beyond what the cluster kernel keeps resident (20 480 atoms for passes that visit every atom twice): the fit and the
    elementwise sweep over the whole frame are separate launches (transform.cu).  chem.Super's in-place result, the per-frame
    outputs and align.RMSDTraj's per-atom sums against the oracle, for all-atom fits and for subset fits (gathered)."""
    nframes = 21
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    idx = np.sort(np.random.default_rng(natoms + nidx).choice(natoms, nidx, replace=False)) if nidx else None
    t = device_traj(frames)
    out = t.super_rmsd(ref, idx, idx, write_back=True)
    moved = t.download()
    work = frames.copy()
    rm, rot, t1, t2 = oracle.super_rmsd_traj(work, ref, idx, write_back=True, nthreads=8)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
    assert np.abs(out["trans1"] - t1).max() < TOL_XYZ and np.abs(out["trans2"] - t2).max() < TOL_XYZ
    assert np.abs(moved - work).max() < TOL_XYZ
    # per-atom sums (align.RMSDTraj), then with the aligned trajectory written back (writeTraj, lovo.go:337-339)
    t.upload(frames)
    got = t.peratom_dev_sum(ref, idx)
    sel = idx if idx is not None else np.arange(natoms)
    want, total, aligned = oracle.rmsd_traj(frames, ref, sel, cpus=7, want_aligned=True)
    assert total == nframes
    assert np.abs(got / nframes - want).max() < 1e-10
    assert np.array_equal(t.download(), frames)       # untouched without write_back
    got = t.peratom_dev_sum(ref, idx, write_back=True)
    assert np.abs(got / nframes - want).max() < 1e-10
    assert np.abs(t.download() - aligned).max() < TOL_XYZ
    t.close()


@pytest.mark.parametrize("natoms,nfit,nwant,nframes", [(9000, 300, 500, 70), (33000, 900, 1500, 33), (700, 0, 40, 129), (5000, 2000, 1, 10)])
def test_per_atom_sums_of_chosen_atoms(natoms, nfit, nwant, nframes):
    """gcb_peratom_dev_sum_atoms: the sums align.RMSDTraj forms, for a chosen list of atoms only (in the caller's order), against
    the full-frame entry point and the oracle; frames stay untouched."""
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    rng = np.random.default_rng(natoms + nwant)
    idx = np.sort(rng.choice(natoms, nfit, replace=False)) if nfit else None
    want = rng.choice(natoms, nwant, replace=False)          # not sorted on purpose
    t = device_traj(frames)
    got = t.peratom_dev_sum_atoms(ref, idx, want)
    full = t.peratom_dev_sum(ref, idx)
    assert np.abs(got - full[want]).max() < 1e-10 * nframes
    sel = idx if idx is not None else np.arange(natoms)
    orc, total = oracle.rmsd_traj(frames, ref, sel, cpus=1)
    assert total == nframes and np.abs(got / nframes - orc[want]).max() < 1e-10
    assert np.array_equal(t.download(), frames)
    with pytest.raises(_lib.GcbError):
        t.peratom_dev_sum_atoms(ref, idx, np.array([0, natoms]))
    t.close()


def test_large_frame_split_path_on_a_window():
    """The split path on a window of the trajectory (first > 0): only those frames move, the per-frame outputs land at the window's
    offsets, and the sums cover the window's frames alone (full-frame and chosen-atoms entry points)."""
    natoms, nframes, first, n = 26000, 12, 3, 7
    ref, frames = synth.make_case(natoms, nframes)
    idx = np.sort(np.random.default_rng(11).choice(natoms, 400, replace=False))
    t = device_traj(frames)
    win = frames[first:first + n].copy()
    want, total = oracle.rmsd_traj(win, ref, idx, cpus=1)
    got = t.peratom_dev_sum(ref, idx, first=first, nframes=n)
    assert total == n and np.abs(got / n - want).max() < 1e-10
    cand = np.array([25999, 0, 17, 13000])
    got_c = t.peratom_dev_sum_atoms(ref, idx, cand, first=first, nframes=n)
    assert np.abs(got_c / n - want[cand]).max() < 1e-10
    out = t.super_rmsd(ref, idx, idx, write_back=True, first=first, nframes=n)
    rm, rot, t1, t2 = oracle.super_rmsd_traj(win, ref, idx, write_back=True)
    assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
    moved = t.download()
    assert np.abs(moved[first:first + n] - win).max() < TOL_XYZ
    assert np.array_equal(moved[:first], frames[:first]) and np.array_equal(moved[first + n:], frames[first + n:])
    # one frame, and an empty window
    one = t.super_rmsd(ref, None, None, write_back=True, first=0, nframes=1)
    assert abs(one["rmsd"][0] - oracle.super_rmsd_traj(frames[:1], ref, None)[0][0]) < TOL_RMSD
    assert np.array_equal(t.peratom_dev_sum(ref, idx, first=2, nframes=0), np.zeros(natoms))
    t.close()


@pytest.mark.parametrize("natoms", [64, 100, 300, 333, 480, 481, 560, 900, 901])
def test_small_structures_split_form(natoms):
    """Write-back and per-atom sums on a few hundred atoms take the split form by default (fit-only launch, then one elementwise
    sweep whose threads own a column and walk the frames): against the oracle and against the fused kernels."""
    lib = _lib.load()
    nframes = 203
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=True)
    idx = np.arange(1, natoms, 3)
    t = device_traj(frames)
    try:
        for sel in (None, idx):
            dev = t.peratom_dev_sum(ref, sel)
            s = sel if sel is not None else np.arange(natoms)
            want, total = oracle.rmsd_traj(frames, ref, s, cpus=7)
            assert total == nframes and np.abs(dev / nframes - want).max() < 1e-10
            t.set_split_thresholds(0, 0)
            fused = t.peratom_dev_sum(ref, sel)
            t.set_split_thresholds(-1, -1)
            assert np.abs(dev - fused).max() < 1e-10 * nframes
            assert np.array_equal(t.download(), frames)
            out = t.super_rmsd(ref, sel, sel, write_back=True)
            work = frames.copy()
            rm, rot, t1, t2 = oracle.super_rmsd_traj(work, ref, sel, write_back=True, nthreads=4)
            assert np.abs(out["rmsd"] - rm).max() < TOL_RMSD and np.abs(out["rot"] - rot).max() < TOL_ROT
            assert np.abs(out["trans1"] - t1).max() < TOL_XYZ and np.abs(out["trans2"] - t2).max() < TOL_XYZ
            assert np.abs(t.download() - work).max() < TOL_XYZ
            t.upload(frames)
    finally:
        t.close()
