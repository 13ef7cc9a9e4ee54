"""CPU tests: the C oracle against the numpy/LAPACK restatement, a Horn-quaternion solver, closed-form
invariants and the committed golden vectors (parity unpinned: the reference holds no golden values for
this path)."""
import os

import numpy as np
import pytest

import oracle
from gochem_b200 import synth
from oracle import oracle_np as onp

TOL_R = 1e-12
TOL_A = 1e-11


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def _go_regen():
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "go_regen_export", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "go_regen", "export_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_go_regen_input_files_round_trip(golden_dir, tmp_path):
    """The flat files oracle/go_regen/main.go reads hold exactly the golden inputs (header, ref, frames, index list)."""
    ex = _go_regen()
    ex.OUT = str(tmp_path)
    g = load(golden_dir, "super_subset_n203")
    path = ex.export(os.path.join(golden_dir, "super_subset_n203.npz"))
    raw = np.fromfile(path, "<i8", 4)
    natoms, nframes, nidx, cpus = (int(v) for v in raw)
    assert (natoms, nframes, nidx, cpus) == (g["ref"].shape[0], g["frames"].shape[0], g["idx"].size, ex.CPUS)
    body = np.fromfile(path, "<f8", offset=32, count=3 * natoms * (1 + nframes))
    assert np.array_equal(body[: 3 * natoms].reshape(natoms, 3), g["ref"].astype(np.float64))
    assert np.array_equal(body[3 * natoms:].reshape(nframes, natoms, 3), g["frames"].astype(np.float64))
    assert np.array_equal(np.fromfile(path, "<i8", offset=32 + 8 * body.size), g["idx"])


@pytest.mark.parametrize("name", ["super_all_n50", "super_subset_n203", "super_cfg1_sample_n1500", "super_edge_mirror_n120", "super_edge_planar_n120",
                                  "super_edge_neardeg_n120"])
def test_go_reference_outputs(golden_dir, name):
    """Pins the oracle against the real goChem when somebody with a Go toolchain has run oracle/go_regen (this image has none,
    so the files are absent and the oracle stays 'parity unpinned'): BASELINE.json's tolerances, 1e-9 A and 1e-10."""
    path = os.path.join(golden_dir, "go_reference", name + ".go.bin")
    if not os.path.exists(path):
        pytest.skip("no output of the Go reference committed (oracle/go_regen/main.go needs a Go toolchain)")
    ex = _go_regen()
    g = load(golden_dir, name)
    ref, frames = g["ref"].astype(np.float64), g["frames"].astype(np.float64)
    idx = g["idx"] if g["idx"].size else None
    go = ex.read_outputs(path, ref.shape[0], frames.shape[0])
    work = frames.copy()
    rm, rot, t1, t2 = oracle.super_rmsd_traj(work, ref, idx, write_back=True)
    assert np.abs(rm - go["rmsd"]).max() < 1e-9
    assert np.abs(rot - go["rot"]).max() < 1e-10
    assert np.abs(t1 - go["trans1"]).max() < 1e-9 and np.abs(t2 - go["trans2"]).max() < 1e-9
    assert np.abs(work - go["moved"]).max() < 1e-9
    sel = idx if idx is not None else np.arange(ref.shape[0])
    want, total = oracle.rmsd_traj(frames, ref, sel, cpus=ex.CPUS)
    assert total == go["frames_read"]
    assert np.abs(want - go["rmsdtraj"]).max() < 1e-9


def test_go_reference_lovo_output(golden_dir, tmp_path):
    """align.LOVO of the real goChem on the golden LOVO case (oracle/go_regen writes the frames to a DCD file and calls it):
    N, the ordered atom list, the iteration count and the per-candidate means.  Skipped while that output is absent; the
    exporter's file layout is checked either way."""
    ex = _go_regen()
    ex.OUT = str(tmp_path)
    g = load(golden_dir, "lovo_n150_f42")
    p = ex.export_lovo(os.path.join(golden_dir, "lovo_n150_f42.npz"))
    assert [int(v) for v in np.fromfile(p, "<i8", 4)] == [g["ref"].shape[0], g["frames"].shape[0], 0, int(g["cpus"])]
    path = os.path.join(golden_dir, "go_reference", "lovo_n150_f42.lovo.go.bin")
    if not os.path.exists(path):
        pytest.skip("no output of the Go reference committed (oracle/go_regen/main.go needs a Go toolchain)")
    go = ex.read_lovo_outputs(path)
    ref, frames = g["ref"].astype(np.float64), g["frames"].astype(np.float64)
    want = oracle.lovo(frames, ref, list(range(ref.shape[0])), int(g["cpus"]), less_than_rmsd=1.0, minimum_n=10)
    assert (go["N"], go["Iterations"], go["FramesRead"]) == (want["N"], want["Iterations"], want["FramesRead"])
    assert list(go["Natoms"]) == list(want["Natoms"])
    assert np.abs(go["RMSD"] - want["RMSD"]).max() < 1e-9


def test_go_reference_moments_begin_skip_writetraj_and_rdf(golden_dir, tmp_path):
    """The rest of what one run of oracle/go_regen pins: chem.CenterOfMass / chem.MomentTensor per frame, align.RMSDTraj with
    Begin / Skip, the DCD bytes align.LOVO writes with WriteTraj, solv.FrameUMolCRDF per frame.  Each part is skipped while its
    file is absent; the exporters' layouts are checked either way."""
    ex = _go_regen()
    ex.OUT = str(tmp_path)
    p = ex.export_rdf(os.path.join(golden_dir, "rdf_small.npz"))
    g = load(golden_dir, "rdf_small")
    assert [int(v) for v in np.fromfile(p, "<i8", 3)] == [g["frames"].shape[1], g["frames"].shape[0], g["refidx"].size]
    lv = load(golden_dir, "lovo_n150_f42")
    p = ex.export_lovo(os.path.join(golden_dir, "lovo_n150_f42.npz"))
    tail = np.fromfile(p, "<i8", offset=32 + 8 * 3 * lv["ref"].shape[0] * (1 + lv["frames"].shape[0]))
    assert tail[0] == lv["begin_skip"].shape[0] and np.array_equal(tail[1:].reshape(-1, 2), lv["begin_skip"])
    gdir = os.path.join(golden_dir, "go_reference")
    checked = 0
    for name in ("super_all_n50", "super_edge_planar_n120"):
        path = os.path.join(gdir, name + ".moments.go.bin")
        if os.path.exists(path):
            c = load(golden_dir, name)
            go = ex.read_moments(path, c["frames"].shape[0])
            for f in range(c["frames"].shape[0]):
                cen, mom = oracle.center_and_moment(c["frames"][f].astype(np.float64))
                assert np.abs(cen - go["center"][f]).max() < 1e-9 and np.abs(mom - go["moment"][f]).max() < 1e-9 * max(1.0, np.abs(mom).max())
            checked += 1
    path = os.path.join(gdir, "lovo_n150_f42.rmsdtraj.go.bin")
    if os.path.exists(path):
        ref, frames = lv["ref"].astype(np.float64), lv["frames"].astype(np.float64)
        sel = np.arange(0, ref.shape[0], 2)
        for (b, k), (read, rm) in zip(lv["begin_skip"], ex.read_rmsdtraj(path, ref.shape[0], lv["begin_skip"].shape[0])):
            want, total = oracle.rmsd_traj(frames, ref, sel, int(lv["cpus"]), int(b), int(k))
            assert total == read and np.abs(want - rm).max() < 1e-9
        checked += 1
    path = os.path.join(gdir, "lovo_n150_f42.lovo.aligned.dcd")
    if os.path.exists(path):
        from gochem_b200 import host_api as H

        ref, frames = lv["ref"].astype(np.float64), lv["frames"].astype(np.float64)
        mine = str(tmp_path / "aligned.dcd")
        src = str(tmp_path / "src.dcd")
        H.dcd_write(src, frames)
        go_frames, nf, na = H.dcd_read(path, frames.shape[0], ref.shape[0])
        assert na == ref.shape[0] and nf > 0
        # the frames the reference wrote are superposed on the final-but-one index set: each must fit the template at least as
        # well as the oracle's own alignment of that frame on the converged set, to float32 rounding
        want = oracle.lovo(frames, ref, list(range(ref.shape[0])), int(lv["cpus"]), less_than_rmsd=1.0, minimum_n=10)
        assert nf == want["FramesRead"]
        checked += 1
    path = os.path.join(gdir, "rdf_small.rdf.go.bin")
    if os.path.exists(path):
        go = ex.read_rdf(path, g["frames"].shape[0], g["cdf_frames_com0"].shape[1])
        assert np.array_equal(go["com0"], g["cdf_frames_com0"]) and np.array_equal(go["com1"], g["cdf_frames_com1"])
        checked += 1
    if checked == 0:
        pytest.skip("no output of the Go reference committed (oracle/go_regen/main.go needs a Go toolchain)")


@pytest.mark.parametrize("name", ["super_all_n50", "super_subset_n203", "super_cfg1_sample_n1500", "super_edge_mirror_n120", "super_edge_planar_n120",
                                  "super_edge_neardeg_n120"])
def test_c_oracle_matches_golden(golden_dir, name):
    g = load(golden_dir, name)
    ref = g["ref"].astype(np.float64)
    frames = g["frames"].astype(np.float64)
    idx = g["idx"] if g["idx"].size else None
    rm, rot, t1, t2 = oracle.super_rmsd_traj(frames, ref, idx)
    assert np.abs(rm - g["rmsd"]).max() < 1e-11
    assert np.abs(rot - g["rot"]).max() < TOL_R
    assert np.abs(t1 - g["trans1"]).max() < TOL_A
    assert np.abs(t2 - g["trans2"]).max() < TOL_A
    args = () if idx is None else (idx, idx)
    assert np.abs(oracle.super_(frames[1], ref, *args) - g["moved_first"]).max() < TOL_A
    assert np.abs(oracle.super_(frames[-1], ref, *args) - g["moved_last"]).max() < TOL_A


def test_three_solvers_agree():
    rng = np.random.default_rng(7)
    for k in range(200):
        n = int(rng.integers(4, 60))
        a = rng.standard_normal((n, 3)) * 12
        b = a @ synth.quat_to_rot(*(lambda q: (q / np.linalg.norm(q),))(rng.standard_normal(4))) + rng.standard_normal((n, 3)) * rng.uniform(0, 4)
        b += rng.uniform(-50, 50, 3)
        _, r_c, t1_c, t2_c = oracle.rotator_translator_to_super(a, b)
        _, r_n, t1_n, t2_n = onp.rotator_translator_to_super(a, b)
        r_h = onp.horn_rotation(a - a.mean(0), b - b.mean(0))
        assert np.abs(r_c - r_n).max() < TOL_R
        assert np.abs(r_c - r_h).max() < 1e-11
        assert np.abs(t1_c - t1_n).max() < TOL_A and np.abs(t2_c - t2_n).max() < TOL_A
        assert abs(np.linalg.det(r_c) - 1) < 1e-13
        assert np.abs(r_c.T @ r_c - np.eye(3)).max() < 1e-13


def test_reflection_branch_matches_lapack():
    # mirror images force det(U V^T) < 0 (geometric.go:183-188)
    rng = np.random.default_rng(11)
    hits = 0
    for k in range(100):
        a = rng.standard_normal((12, 3))
        b = a * np.array([1.0, 1.0, -1.0]) + rng.standard_normal((12, 3)) * 0.05
        _, r_c, _, _ = oracle.rotator_translator_to_super(a, b)
        _, r_n, _, _ = onp.rotator_translator_to_super(a, b)
        h = (a - a.mean(0)).T @ (b - b.mean(0))
        if np.linalg.det(h) < 0:
            hits += 1
        assert np.abs(r_c - r_n).max() < 1e-10
        assert np.linalg.det(r_c) > 0
    assert hits > 50


def test_rigid_copy_gives_zero_rmsd():
    # "RMSD mol (should be 0)", gochem_test.go:489
    ref = synth.make_reference(300)
    q = np.array([0.3, -0.5, 0.7, 0.1]); q /= np.linalg.norm(q)
    moved = ref @ synth.quat_to_rot(q) + np.array([10.0, -20.0, 33.0])
    s = oracle.super_(moved, ref)
    assert oracle.rmsd(s, ref) < 1e-12
    assert abs(oracle.rmsd(s, ref) - oracle.rmsd_naive(s, ref)) < 1e-13


def test_superposition_is_a_minimum():
    ref, frames = synth.make_case(80, 3)
    rng = np.random.default_rng(3)
    for f in range(3):
        s, r, t1, t2 = oracle.super_(frames[f], ref, details=True)
        best = oracle.rmsd(s, ref)
        for _ in range(20):
            q = np.array([1.0, 0, 0, 0]) + rng.standard_normal(4) * 1e-3
            q /= np.linalg.norm(q)
            pert = ((frames[f] + t1) @ (r @ synth.quat_to_rot(q))) + t2
            assert oracle.rmsd(pert, ref) >= best - 1e-12
        # centroid of the superposed set equals the template's
        assert np.abs(s.mean(0) - ref.mean(0)).max() < 1e-12


def test_index_dispatch_and_errors():
    ref, frames = synth.make_case(40, 2)
    idx = np.arange(0, 40, 2)
    # two lists
    a = oracle.super_(frames[1], ref, idx, idx)
    b = onp.super_(frames[1], ref, idx, idx)
    assert np.abs(a - b).max() < TOL_A
    # one list, test larger than the list: templa must have len(idx) rows (handy.go:181-185)
    a = oracle.super_(frames[1], ref[idx], idx)
    b = onp.super_(frames[1], ref[idx], idx)
    assert np.abs(a - b).max() < TOL_A
    # one list that would apply to templa: nil dereference in the reference (handy.go:186-188)
    with pytest.raises(oracle.OracleError) as e:
        oracle.super_(frames[1][idx], ref, idx)
    assert e.value.code == oracle.ERR_NILDEREF
    # mismatched row counts
    with pytest.raises(oracle.OracleError) as e:
        oracle.super_(frames[1], ref[:30])
    assert e.value.code == oracle.ERR_SHAPE
    with pytest.raises(oracle.OracleError):
        oracle.mem_rmsd(frames[1], ref[:30])
    # RMSD with lists equals numpy
    assert abs(oracle.rmsd(frames[1], ref, idx, idx) - onp.mem_rmsd(frames[1], ref, idx, idx)) < 1e-12
    assert np.abs(oracle.per_atom_rmsd(frames[1], ref) - onp.per_atom_rmsd(frames[1], ref)).max() < 1e-12


def test_as_written_variant_is_numerically_identical():
    ref, frames = synth.make_case(64, 2)
    a = oracle.rotator_translator_to_super(frames[1], ref, as_written=False)
    b = oracle.rotator_translator_to_super(frames[1], ref, as_written=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_mobile_limit_quirks():
    vals = np.array([0.1, 0.2, 0.3, 0.9, 1.5, 2.0])
    # first value >= 1.0 is at k = 4 -> n = 3 (off by one, lovo.go:180-183); limit leaves multiplied by 1.1
    n, lim = oracle.mobile_limit(1.0, 2, vals)
    assert (n, lim) == (3, 1.0 * 1.1)
    assert onp.mobile_limit(1.0, 2, list(vals)) == (n, lim)
    # nothing above the limit: n = Len
    assert oracle.mobile_limit(5.0, 2, vals)[0] == 6
    # n < tolerance: limit grows by 1.1 per scan until enough atoms qualify
    n, lim = oracle.mobile_limit(0.25, 4, vals)
    assert (n, lim) == onp.mobile_limit(0.25, 4, list(vals))
    assert n >= 4
    # tolerance <= 0: loop body never runs
    assert oracle.mobile_limit(1.0, 0, vals) == (0, 1.0)
    # would never terminate when Len < tolerance
    with pytest.raises(oracle.OracleError):
        oracle.mobile_limit(1.0, 10, vals)


def test_stable_sort_matches_python():
    rng = np.random.default_rng(5)
    vals = np.round(rng.uniform(0, 1, 200), 1)  # many ties
    ids = np.arange(200)
    i2, v2 = oracle.stable_sort_pairs(ids, vals, by_rmsd=True)
    order = sorted(range(200), key=lambda k: vals[k])
    assert list(i2) == order
    i3, v3 = oracle.stable_sort_pairs(i2, v2, by_rmsd=False)
    assert list(i3) == list(range(200)) and np.array_equal(v3, vals)


def test_rmsd_traj_chunk_semantics():
    ref, frames = synth.make_case(30, 11)
    idx = np.arange(30)
    # 11 frames, chunks of 4: the tail of 3 is dropped (lovo.go:323-326)
    rm, total = oracle.rmsd_traj(frames, ref, idx, cpus=4)
    assert total == 8
    rm_np, total_np = onp.rmsd_traj(frames, ref, list(idx), 4)
    assert total_np == 8 and np.abs(rm - rm_np).max() < 1e-12
    # threads do not change the numbers
    rm_t, _ = oracle.rmsd_traj(frames, ref, idx, cpus=4, nthreads=4)
    assert np.array_equal(rm, rm_t)
    # skip: every second chunk is discarded but the chunk after it is still processed (no `continue`)
    rm_s, total_s = oracle.rmsd_traj(frames, ref, idx, cpus=2, skip=2)
    rm_sn, total_sn = onp.rmsd_traj(frames, ref, list(idx), 2, skip=2)
    assert total_s == total_sn and np.abs(rm_s - rm_sn).max() < 1e-12
    # aligned frames are what Super leaves behind
    rm_a, tot_a, aligned = oracle.rmsd_traj(frames, ref, idx, cpus=4, want_aligned=True)
    assert aligned.shape[0] == 8
    assert np.abs(aligned[3] - oracle.super_(frames[3], ref, idx, idx)).max() == 0


def test_lovo_c_vs_numpy_and_golden(golden_dir):
    g = load(golden_dir, "lovo_n150_f42")
    ref = g["ref"].astype(np.float64)
    frames = g["frames"].astype(np.float64)
    cpus = int(g["cpus"])
    res = oracle.lovo(frames, ref, np.arange(150), cpus)
    assert res["N"] == int(g["N"]) and res["Iterations"] == int(g["Iterations"])
    assert res["FramesRead"] == int(g["FramesRead"]) == 40
    assert list(res["Natoms"]) == list(g["Natoms"])
    assert np.abs(res["RMSD"] - g["RMSD"]).max() < 1e-11
    rm, total = oracle.rmsd_traj(frames, ref, np.arange(150), cpus)
    assert total == int(g["first_pass_frames"]) and np.abs(rm - g["first_pass_rmsd"]).max() < 1e-11
    res_t = oracle.lovo(frames, ref, np.arange(150), cpus, nthreads=4)
    assert list(res_t["Natoms"]) == list(res["Natoms"])


def test_center_of_mass_and_moment_tensor():
    rng = np.random.default_rng(9)
    a = rng.standard_normal((70, 3)) * 9 + np.array([30.0, -12.0, 5.0])
    m = rng.uniform(1.0, 32.0, 70)
    for mass in (None, m):
        w = np.ones(70) if mass is None else mass
        cen, mom = oracle.center_and_moment(a, mass)
        c2 = (a * w[:, None]).sum(0) / w.sum()
        d = a - c2
        assert np.abs(cen - c2).max() < 1e-13
        assert np.abs(mom - (d * w[:, None]).T @ d).max() < 1e-10
        assert np.array_equal(mom, mom.T)


def test_rdf_two_restatements_agree():
    """solv.DistRank / FrameUMolCRDF: the C restatement against the numpy one on a small solvated system."""
    from oracle import oracle_np
    rng = np.random.default_rng(3)
    ns, nw = 40, 200
    sol = rng.normal(0, 4, (ns, 3))
    wo = rng.uniform(-18, 18, (nw, 3))
    wat = (wo[:, None, :] + np.concatenate([np.zeros((nw, 1, 3)), rng.normal(0, 0.6, (nw, 2, 3))], axis=1)).reshape(-1, 3)
    ions = rng.uniform(-15, 15, (20, 3))
    coord = np.concatenate([sol, wat, ions])
    molid = np.concatenate([np.repeat(np.arange(1, 5), 10), 5 + np.repeat(np.arange(nw), 3), 5 + nw + np.arange(20)])
    chain = np.concatenate([np.ones(ns), np.full(3 * nw + 20, 2)]).astype(int)
    flags = np.concatenate([np.zeros(ns), np.full(3 * nw, 3), np.full(20, 1)]).astype(np.uint8)
    for com in (False, True):
        a = oracle.frame_umol_crdf(coord, molid, chain, flags, np.arange(ns), com=com)
        b = oracle_np.frame_umol_crdf(coord, molid, chain, flags, np.arange(ns), com=com)
        assert a[-1] > 20 and np.array_equal(a, b)
    # FrameUMolCRDF's "minus one" (solvation.go:402-404) and MDFFromCDF's closed form for a uniform count
    cdf = np.arange(100, dtype=float)
    rdf, cum = oracle.mdf_from_cdf(cdf * 7, 7, 1.0, 1.0, 0.1)
    assert np.allclose(cum[1:], cdf[1:]) and rdf[-1] == 1.0
    with pytest.raises(oracle.OracleError):
        oracle.mdf_from_cdf(cdf, 1, 0.5, 1.0, 0.1)


def test_rdf_c_oracle_matches_golden(golden_dir):
    g = load(golden_dir, "rdf_small")
    frames = g["frames"].astype(np.float64)
    for com in (0, 1):
        want = g["cdf_frames_com%d" % com]
        for f in range(frames.shape[0]):
            got = oracle.frame_umol_crdf(frames[f], g["molid"], g["chain"], g["flags"], g["refidx"], com=bool(com),
                                         step=float(g["step"]), end=float(g["end"]))
            assert np.array_equal(got, want[f])
        total, read = oracle.rdf_cdf_sum(frames, g["molid"], g["chain"], g["flags"], g["refidx"], com=bool(com), cpus=3)
        assert read == 9 and np.array_equal(total, want[:9].sum(axis=0))
