"""GPU tests of the host-side mirror of the reference's Go API (chem.Super, chem.RMSD, align.RMSDTraj,
align.LOVO) against the oracle.  They read like the reference's own tests (gochem_test.go:451-495,
align/lovo_test.go:34-59), except that they assert."""
import os

import numpy as np
import pytest

import oracle
from gochem_b200 import _lib
from gochem_b200 import host_api as H
from gochem_b200 import synth

pytestmark = pytest.mark.gpu


def test_super_and_rmsd_like_TTestSuper():
    ref, frames = synth.make_case(400, 3)
    test = frames[1]
    idx = np.arange(5, 395, 3)
    # no lists
    s = H.Super(test, ref)
    assert np.abs(s - oracle.super_(test, ref)).max() < 1e-9
    assert abs(H.RMSD(s, ref) - oracle.rmsd(oracle.super_(test, ref), ref)) < 1e-9
    assert abs(H.RMSD(s, ref) - oracle.rmsd_naive(s, ref)) < 1e-9   # the cross-check of gochem_test.go:472-473
    # two lists
    s2 = H.Super(test, ref, idx, idx)
    assert np.abs(s2 - oracle.super_(test, ref, idx, idx)).max() < 1e-9
    assert abs(H.RMSD(s2, ref, idx, idx) - oracle.rmsd(s2, ref, idx, idx)) < 1e-9
    # one list: applies to test, templa holds len(list) rows (handy.go:181-185)
    s3 = H.Super(test, ref[idx], idx)
    assert np.abs(s3 - oracle.super_(test, ref[idx], idx)).max() < 1e-9
    # "RMSD mol (should be 0)" (gochem_test.go:489): a rigidly moved copy
    q = np.array([0.2, 0.4, -0.3, 0.8]); q /= np.linalg.norm(q)
    moved = ref @ synth.quat_to_rot(q) + np.array([3.0, -40.0, 12.5])
    assert H.RMSD(H.Super(moved, ref), ref) < 1e-12
    # per-atom distances and the low-level call
    assert np.abs(H.MemPerAtomRMSD(s, ref) - oracle.per_atom_rmsd(s, ref)).max() < 1e-12
    assert np.abs(H.MemPerAtomRMSD(s, ref, idx, idx) - oracle.per_atom_rmsd(s[idx], ref[idx])).max() < 1e-12
    tr, r, a, b = H.RotatorTranslatorToSuper(test, ref)
    otr, orr, oa, ob = oracle.rotator_translator_to_super(test, ref)
    assert np.abs(tr - otr).max() < 1e-9 and np.abs(r - orr).max() < 1e-10
    assert np.abs(a - oa).max() < 1e-9 and np.abs(b - ob).max() < 1e-9


def test_errors_read_like_the_reference():
    ref, frames = synth.make_case(30, 2)
    idx = np.arange(0, 30, 2)
    with pytest.raises(H.GoChemError, match="Ill formed coordinates"):
        H.Super(frames[1], ref[:20])
    with pytest.raises(H.GoChemError, match="Ill formed coordinates"):
        H.Super(frames[1], ref, idx, idx[:5])
    with pytest.raises(H.GoChemError, match="nil"):
        H.Super(frames[1][idx], ref, idx)       # single list that would apply to templa (handy.go:186-188)
    with pytest.raises(H.GoChemError, match="Indexes don't match"):
        H.Super(frames[1][idx], ref[idx], idx)
    with pytest.raises(H.GoChemError, match="Ill formed"):
        H.RMSD(frames[1], ref[:20])
    with pytest.raises(H.GoChemError, match="Ill-formed"):
        H.RotatorTranslatorToSuper(frames[1], ref[:20])


def test_rmsdtraj_on_a_dcd_file(tmp_path):
    ref, frames = synth.make_case(256, 23, through_f32=True)
    p = str(tmp_path / "traj.dcd")
    H.dcd_write(p, frames)
    idx = np.arange(0, 256, 2)
    for cpus, begin, skip in [(4, 0, 0), (3, 0, 0), (2, 4, 0), (2, 0, 2), (5, 5, 5)]:
        got, nframes = H.RMSDTraj(p, ref, idx, cpus, begin, skip)
        want, total = oracle.rmsd_traj(frames, ref, idx, cpus, begin, skip)
        assert nframes == total
        assert np.abs(got - want).max() < 1e-10
    # writeTraj: the aligned frames end up in a DCD file (lovo.go:296-305, 337-339), float32 like the reference writes
    out = str(tmp_path / "aligned.dcd")
    got, nframes = H.RMSDTraj(p, ref, idx, 4, write_traj=out)
    want, total, aligned = oracle.rmsd_traj(frames, ref, idx, 4, want_aligned=True)
    back, nf, na = H.dcd_read(out, 30, 256)
    assert nf == total == 20
    assert np.abs(back - aligned.astype(np.float32)).max() < 1e-4
    # no such file (or no libxdrfile on this host): an error, as opentraj reports it (lovo.go:157-160)
    with pytest.raises(H.GoChemError, match="libxdrfile|Probem opening"):
        H.RMSDTraj(str(tmp_path / "x.xtc"), ref, idx, 4)
    with pytest.raises(H.GoChemError, match="no such file|libzstd"):      # traj/stf is opened like the others (lovo.go:161-162)
        H.RMSDTraj(str(tmp_path / "x.stf"), ref, idx, 4)
    with pytest.raises(H.GoChemError, match="not supported|outside this build"):
        H.RMSDTraj(str(tmp_path / "x.pdb"), ref, idx, 4)
    with pytest.raises(H.GoChemError, match="not supported"):
        H.RMSDTraj(str(tmp_path / "x.trr"), ref, idx, 4)


def test_sharded_rmsdtraj_partials_add_up(tmp_path):
    ref, frames = synth.make_case(128, 41, through_f32=True)
    p = str(tmp_path / "traj.dcd")
    H.dcd_write(p, frames)
    idx = np.arange(128)
    whole, n = H.RMSDTraj(p, ref, idx, 4)
    parts = [H.RMSDTraj(p, ref, idx, 4, rank=r, nranks=3)[0] for r in range(3)]   # no communicator: local share / total
    assert np.abs(sum(parts) - whole).max() < 1e-12


def test_streamed_rmsdtraj_and_lovo_match_the_resident_path(tmp_path):
    """Bounded-memory variants: the DCD file goes through a device window of a few frames (several windows, a ragged last one,
    windows dealt to ranks in turn); same chunk rules, same results as the resident path and the oracle."""
    ref, frames = synth.make_case(300, 61, through_f32=True)
    p = str(tmp_path / "traj.dcd")
    H.dcd_write(p, frames)
    idx = np.arange(0, 300, 3)
    for cpus, begin, skip, window in [(4, 0, 0, 8), (3, 0, 0, 7), (2, 4, 2, 2), (5, 0, 0, 1000)]:
        got, nframes = H.RMSDTrajStreamed(p, ref, idx, cpus, window, begin, skip)
        want, total = oracle.rmsd_traj(frames, ref, idx, cpus, begin, skip)
        res, nres = H.RMSDTraj(p, ref, idx, cpus, begin, skip)
        assert nframes == total == nres
        assert np.abs(got - want).max() < 1e-10 and np.abs(got - res).max() < 1e-12
    # ranks take windows in turn; without a communicator each returns its share / total
    parts = [H.RMSDTrajStreamed(p, ref, idx, 4, 8, rank=r, nranks=3)[0] for r in range(3)]
    whole, _ = H.RMSDTrajStreamed(p, ref, idx, 4, 8)
    assert np.abs(sum(parts) - whole).max() < 1e-12
    # writeTraj while streaming: the aligned frames in reading order
    out = str(tmp_path / "aligned_stream.dcd")
    H.RMSDTrajStreamed(p, ref, idx, 4, 8, write_traj=out)
    _, total, aligned = oracle.rmsd_traj(frames, ref, idx, 4, want_aligned=True)
    back, nf, na = H.dcd_read(out, 80, 300)
    assert nf == total == 60 and np.abs(back - aligned.astype(np.float32)).max() < 1e-4
    # LOVO re-reading the file on every iteration (lovo.go:222) == LOVO on the resident trajectory
    a = H.LOVOStreamed(p, ref, 4, 12)
    b = H.LOVO(p, ref, 4)
    assert a["N"] == b["N"] and a["Iterations"] == b["Iterations"] and list(a["Natoms"]) == list(b["Natoms"])
    assert a["FramesRead"] == b["FramesRead"] and np.abs(a["RMSD"] - b["RMSD"]).max() < 1e-12


def test_lovo_like_TestLovo(tmp_path, golden_dir):
    z = np.load(os.path.join(golden_dir, "lovo_n150_f42.npz"))
    ref, frames, cpus = z["ref"].astype(np.float64), z["frames"].astype(np.float64), int(z["cpus"])
    p = str(tmp_path / "test_align.dcd")
    H.dcd_write(p, frames)
    got = H.LOVO(p, ref, cpus, less_than_rmsd=1.0)
    assert got["N"] == int(z["N"]) and got["Iterations"] == int(z["Iterations"]) and got["FramesRead"] == 40
    assert list(got["Natoms"]) == list(z["Natoms"])            # ordered list identical, not just the set
    assert np.abs(got["RMSD"] - z["RMSD"]).max() < 1e-10
    assert got["PyMOLSel"].startswith("select rigid, (resi ") and " and chain A" in got["PyMOLSel"]
    assert got["GMX"].startswith("r ") and got["GMX"].endswith("& chain A")
    assert got["String"].startswith(f"N: {got['N']},  Frames read: 40, Iterations needed: {got['Iterations']}, RMSD: [")


def test_lovo_larger_case_and_topology_filter(tmp_path):
    natoms, nframes, cpus = 900, 160, 8
    ref, frames = synth.make_case(natoms, nframes, through_f32=True)
    p = str(tmp_path / "big.dcd")
    H.dcd_write(p, frames)
    # every third atom is a CA of chain A; the rest are side-chain atoms
    names = ["CA" if i % 3 == 0 else "CB" for i in range(natoms)]
    chains = ["A"] * natoms
    molids = [i // 3 + 1 for i in range(natoms)]
    full = np.arange(0, natoms, 3)
    want = oracle.lovo(frames, ref, full, cpus, less_than_rmsd=1.0, minimum_n=10, nthreads=8)
    got = H.LOVO(p, ref, cpus, names=names, chains=chains, molids=molids)
    assert got["N"] == want["N"] and got["Iterations"] == want["Iterations"] and got["FramesRead"] == want["FramesRead"]
    assert list(got["Natoms"]) == list(want["Natoms"])
    assert np.abs(got["RMSD"] - want["RMSD"]).max() < 1e-10
    # the aligned trajectory is written once the index set differs by a single atom (lovo.go:226-228); whether that
    # happens depends on the data, so only check that asking for it does not change the answer
    got2 = H.LOVO(p, ref, cpus, names=names, chains=chains, molids=molids, write_traj=str(tmp_path / "al.dcd"))
    assert list(got2["Natoms"]) == list(want["Natoms"])


def test_rmsdtraj_on_an_xtc_file(tmp_path):
    """traj/xtc through the codec bound with dlopen (libxdrfile where it is installed, tests/xdrshim elsewhere -- conftest.py):
    frames decoded straight into pinned memory as float32 nanometres, x10 / widened / transposed on the device."""
    assert H.xtc_available(), "neither libxdrfile nor tests/xdrshim could be bound"
    ref, frames = synth.make_case(200, 26, through_f32=True)
    p = str(tmp_path / "traj.xtc")
    H.xtc_write(p, frames)
    decoded, nf, na = H.xtc_read(p, 40, 200)          # what the codec hands back (lossy: 0.001 nm)
    assert (nf, na) == (26, 200) and np.abs(decoded - frames).max() < 0.006
    idx = np.arange(0, 200, 2)
    for cpus, window in [(4, 0), (3, 0), (4, 8)]:
        got, n = (H.RMSDTrajStreamed(p, ref, idx, cpus, window) if window else H.RMSDTraj(p, ref, idx, cpus))
        want, total = oracle.rmsd_traj(decoded, ref, idx, cpus)
        assert n == total and np.abs(got - want).max() < 1e-9


@pytest.mark.parametrize("ext", ["stf"])      # opentraj knows the extension "stf" only (lovo.go:161), i.e. the zstd form
def test_rmsdtraj_and_lovo_on_an_stf_file(tmp_path, ext):
    """traj/stf: the integers of the text lines go to the device as parsed (GCB_I32_AOS) and are divided by 100 there; the frames
    the device sees equal float64(f) / 100 of stf.go:341 bit for bit, so RMSDTraj on the file matches the oracle on the decoded
    frames.  zstd through libzstd (skipped where it does not load); the gzip / deflate forms are covered on the CPU
    (tests/test_host_logic.py)."""
    if ext == "stf" and not H.stf_zstd_available():
        pytest.skip("libzstd is not loadable on this host")
    ref, frames = synth.make_case(150, 42)
    p = str(tmp_path / ("traj." + ext))
    H.stf_write(p, frames, level=6)
    decoded, nf, na, _ = H.stf_read(p, 50, 150)
    assert (nf, na) == (42, 150) and np.array_equal(decoded, np.rint(frames * 100.0) / 100.0)
    idx = np.arange(0, 150, 2)
    for cpus, window in [(4, 0), (5, 0), (4, 8)]:
        got, n = (H.RMSDTrajStreamed(p, ref, idx, cpus, window) if window else H.RMSDTraj(p, ref, idx, cpus))
        want, total = oracle.rmsd_traj(decoded, ref, idx, cpus)
        assert n == total and np.abs(got - want).max() < 1e-10
    got = H.LOVO(p, ref, 6)
    want = oracle.lovo(decoded, ref, list(range(150)), 6, less_than_rmsd=1.0, minimum_n=10)
    assert got["N"] == want["N"] and list(got["Natoms"]) == list(want["Natoms"]) and got["Iterations"] == want["Iterations"]


def test_lovo_on_a_solvated_system_reads_the_candidates_only(tmp_path):
    """A protein in its solvent box: one atom in ten is a CA, the rest never enter LOVO's bookkeeping (rmsdFilter,
    lovo.go:519-528).  The host mirror then asks the device for the candidates' sums alone (gcb_peratom_dev_sum_atoms);
    N, the ordered atom list, the iteration count and the candidates' mean distances equal the oracle's."""
    natoms, nframes, cpus = 2000, 96, 8
    ref, frames = synth.make_case(natoms, nframes, through_f32=True)
    p = str(tmp_path / "solvated.dcd")
    H.dcd_write(p, frames)
    names = ["CA" if i % 10 == 0 else "OW" for i in range(natoms)]
    chains = ["A"] * natoms
    molids = [i // 10 + 1 for i in range(natoms)]
    full = np.arange(0, natoms, 10)
    want = oracle.lovo(frames, ref, full, cpus, less_than_rmsd=1.0, minimum_n=10, nthreads=8)
    got = H.LOVO(p, ref, cpus, names=names, chains=chains, molids=molids)
    assert got["N"] == want["N"] and got["Iterations"] == want["Iterations"] and got["FramesRead"] == want["FramesRead"]
    assert list(got["Natoms"]) == list(want["Natoms"])
    assert np.abs(got["RMSD"] - want["RMSD"]).max() < 1e-10


def test_easy_shape_single_and_batched():
    """chem.EasyShape (handy.go:103-129) of one structure and of every resident frame (moment tensors from the device sweep,
    eigenvalues on the host): against numpy on M = sum m (x - c)(x - c)^T, c the centre of mass (geometric.go:670-733)."""
    natoms, nframes = 700, 37
    ref, frames = synth.make_case(natoms, nframes)
    frames[:, :, 0] *= 2.5                       # make the ellipsoid clearly prolate
    mass = np.linspace(1.0, 16.0, natoms)

    def shape(x, m):
        c = (x * m[:, None]).sum(0) / m.sum()
        d = (x - c) * np.sqrt(m)[:, None]
        ev = np.sort(np.linalg.eigvalsh(d.T @ d))[::-1]
        return (1 - ev[1] / ev[0]) * 100, (1 - ev[2] / ev[1]) * 100

    for m in (None, mass):
        lin, circ = H.EasyShape(frames[3], m)
        wl, wc = shape(frames[3], np.ones(natoms) if m is None else m)
        assert abs(lin - wl) < 1e-8 and abs(circ - wc) < 1e-8
        t = _lib.DeviceTraj(natoms, nframes)
        t.upload(frames)
        L, Cc = H.EasyShapeTraj(t, m)
        for f in range(nframes):
            wl, wc = shape(frames[f], np.ones(natoms) if m is None else m)
            assert abs(L[f] - wl) < 1e-8 and abs(Cc[f] - wc) < 1e-8
        L2, _ = H.EasyShapeTraj(t, m, first=5, nframes=4)
        assert np.array_equal(L2, L[5:9])
        t.close()
    # every atom on one point: the reference's "collapsed" error, -1 / -1
    with pytest.raises(H.GoChemError, match="colapsed"):
        H.EasyShape(np.zeros((10, 3)))
