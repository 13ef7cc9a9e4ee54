"""solv.ConcMolRDF's frame loop on the device against the oracle (integers: exact)."""
import os

import numpy as np
import pytest

import oracle
from gochem_b200 import _lib
from gochem_b200 import host_api as H

pytestmark = pytest.mark.gpu


def solvated_system(nsolute=40, nwater=300, nions=20, nframes=24, seed=3, box=18.0):
    """A compact solute (4 residues, chain A), 3-atom waters (SOL) and 1-atom ions (NA), chain B; every frame jiggles everything
    and moves the waters.  Coordinates pass through float32 so a DCD round trip is lossless."""
    rng = np.random.default_rng(seed)
    sol = rng.normal(0, 4, (nsolute, 3))
    wo = rng.uniform(-box, box, (nwater, 3))
    wat = (wo[:, None, :] + np.concatenate([np.zeros((nwater, 1, 3)), rng.normal(0, 0.6, (nwater, 2, 3))], axis=1)).reshape(-1, 3)
    ions = rng.uniform(-box, box, (nions, 3))
    base = np.concatenate([sol, wat, ions])
    n = base.shape[0]
    frames = np.empty((nframes, n, 3))
    for f in range(nframes):
        fr = base + rng.normal(0, 0.3, (n, 3))
        fr[nsolute:nsolute + 3 * nwater] += np.repeat(rng.normal(0, 1.5, (nwater, 3)), 3, axis=0)
        frames[f] = fr + rng.uniform(-5, 5, 3)
    frames = frames.astype(np.float32).astype(np.float64)
    molid = np.concatenate([np.repeat(np.arange(1, 5), nsolute // 4), 5 + np.repeat(np.arange(nwater), 3), 5 + nwater + np.arange(nions)])
    chain = np.concatenate([np.ones(nsolute), np.full(3 * nwater + nions, 2)]).astype(np.int32)
    flags = np.concatenate([np.zeros(nsolute), np.full(3 * nwater, 3), np.full(nions, 1)]).astype(np.uint8)
    names = ["ALA"] * nsolute + ["SOL"] * (3 * nwater) + ["NA"] * nions
    chains = ["A"] * nsolute + ["B"] * (3 * nwater + nions)
    return frames, molid.astype(np.int32), chain, flags, np.arange(nsolute), names, chains


def device_traj(frames):
    t = _lib.DeviceTraj(frames.shape[1], frames.shape[0])
    t.upload(frames)
    return t


@pytest.mark.parametrize("com", [False, True])
@pytest.mark.parametrize("cpus", [1, 5])
def test_rdf_cdf_sum_matches_oracle(com, cpus):
    frames, molid, chain, flags, refidx, _, _ = solvated_system()
    t = device_traj(frames)
    got, read = t.rdf_cdf_sum(molid, chain, flags, refidx, com=com, cpus=cpus)
    want, wread = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, com=com, cpus=cpus)
    t.close()
    assert read == wread == (frames.shape[0] // cpus) * cpus     # the tail a chunk cannot fill is not read
    assert want.max() > 10                                        # the case is not empty
    assert np.array_equal(got, want)


def test_rdf_large_solute_tiles_and_other_shells():
    """More solute atoms than one shared-memory tile, a different shell width / cutoff, a solute given out of order."""
    frames, molid, chain, flags, refidx, _, _ = solvated_system(nsolute=1400, nwater=500, nions=7, nframes=6, seed=11, box=30.0)
    refidx = refidx[::-1].copy()
    t = device_traj(frames)
    for step, end in [(0.25, 7.0), (0.1, 10.0), (1.0, 3.0)]:
        got, read = t.rdf_cdf_sum(molid, chain, flags, refidx, step=step, end=end)
        want, _ = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, step=step, end=end)
        assert got.shape == want.shape and np.array_equal(got, want)
    t.close()


def test_rdf_prescreen_quirk_and_own_residues():
    """A water whose first atom is beyond end + 3 of the first solute atom starts at its next atom (solvation.go:480-486); the
    solute's own residues never count even when their name is asked for."""
    frames, molid, chain, flags, refidx, _, _ = solvated_system(nsolute=8, nwater=60, nions=0, nframes=4, seed=5, box=9.0)
    flags[:8] = 1                                  # solute residue name also in `residues`
    # stretch one water so that its O is far from solute atom 0 and an H is close to the solute
    for f in range(frames.shape[0]):
        a0 = frames[f, 0]
        frames[f, 8] = a0 + np.array([14.5, 0.0, 0.0])      # O: beyond end + 3 = 13
        frames[f, 9] = a0 + np.array([2.0, 0.5, 0.0])       # H: close
        frames[f, 10] = a0 + np.array([13.5, 0.0, 0.0])
    t = device_traj(frames)
    got, _ = t.rdf_cdf_sum(molid, chain, flags, refidx)
    want, _ = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx)
    t.close()
    assert np.array_equal(got, want)
    d, ids = oracle.dist_rank(frames[0], molid, chain, flags, refidx)
    assert 5 in ids and not set(ids) & {1, 2, 3, 4}


def test_rdf_repeated_residue_ids_follow_the_reference():
    """A water that reuses another water's MolID (same chain) used to be refused; the reference handles it with its
    molid_skip / chain_skip rule, and so does the kernel now -- equal to the oracle in both Options.COM modes."""
    frames, molid, chain, flags, refidx, _, _ = solvated_system(nwater=30, nions=0, nframes=4)
    molid = molid.copy()
    molid[-3:] = molid[40]                           # the last water reuses the first water's MolID
    t = device_traj(frames)
    for com in (False, True):
        got, _ = t.rdf_cdf_sum(molid, chain, flags, refidx, com=com)
        want, _ = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, com=com)
        assert np.array_equal(got, want)
    t.close()


def test_conc_mol_rdf_host_mirror_on_a_dcd(tmp_path):
    """solv.ConcMolRDF end to end: DCD file -> device -> MDFFromCDF, against the oracle's frame loop + MDFFromCDF."""
    frames, molid, chain, flags, refidx, names, chains = solvated_system(nframes=22)
    path = os.path.join(tmp_path, "solv.dcd")
    H.dcd_write(path, frames)
    masses = np.random.default_rng(0).uniform(1.0, 16.0, frames.shape[1])
    cpus = 4
    rdf, cum, read = H.ConcMolRDF(path, frames[0], names, molid, chains, masses, refidx, ["SOL", "NA"], cpus=cpus)
    cdf, wread = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, cpus=cpus)
    assert read == wread == 20
    _, mom = oracle.center_and_moment(frames[0][refidx], masses[refidx])
    ev = np.sort(np.linalg.eigvalsh(mom))[::-1]
    want_rdf, want_cum = oracle.mdf_from_cdf(cdf, wread, ev[0] / ev[2], ev[1] / ev[2], 0.1)
    assert np.allclose(cum, want_cum, rtol=0, atol=0)
    assert np.abs(rdf - want_rdf).max() <= 1e-10 * np.abs(want_rdf).max()
    rho = H.EllipsoidAxes(frames[0][refidx], masses[refidx])
    assert np.abs(rho - ev).max() < 1e-9 * ev[0]
    # the same file streamed through a device window of a few frames: identical counts, identical result
    rdf_s, cum_s, read_s = H.ConcMolRDF(path, frames[0], names, molid, chains, masses, refidx, ["SOL", "NA"], cpus=cpus, window_frames=8)
    assert read_s == read and np.array_equal(cum_s, cum) and np.array_equal(rdf_s, rdf)
    # missing masses are an error, as Topology.Masses reports (chem.go:291-301)
    with pytest.raises(H.GoChemError):
        H.ConcMolRDF(path, frames[0], names, molid, chains, None, refidx, ["SOL"], cpus=cpus)


@pytest.mark.parametrize("com", [0, 1])
def test_rdf_golden_fixture(golden_dir, com):
    z = np.load(os.path.join(golden_dir, "rdf_small.npz"))
    frames = z["frames"].astype(np.float64)
    t = device_traj(frames)
    got, read = t.rdf_cdf_sum(z["molid"], z["chain"], z["flags"], z["refidx"], com=bool(com), step=float(z["step"]), end=float(z["end"]))
    assert read == frames.shape[0] and np.array_equal(got, z["cdf_frames_com%d" % com].sum(axis=0))
    # frame by frame through the window arguments
    one, _ = t.rdf_cdf_sum(z["molid"], z["chain"], z["flags"], z["refidx"], com=bool(com), first=4, nframes=1)
    assert np.array_equal(one, z["cdf_frames_com%d" % com][4])
    t.close()


def test_rdf_residue_numbers_that_wrap():
    """A solvated box whose water numbering wraps (resSeq is four digits in PDB files) and whose chain is blank: the same
    (MolID, chain) pair names several waters.  The reference only passes over a residue whose pair equals that of the residue it
    started LAST (molid_skip / chain_skip, solvation.go:486, 537-538); the counts must equal the oracle's, which walks the atoms
    one by one like the reference."""
    frames, molid, chain, flags, refidx, _, _ = solvated_system(nsolute=40, nwater=10500, nions=30, nframes=3, seed=21, box=60.0)
    nsol = 40
    wat = np.arange(10500)
    molid[nsol:nsol + 3 * 10500] = np.repeat(5 + wat % 9999, 3)            # 10500 waters numbered 5 .. 10003, then 5 .. again
    chain[:] = 0                                                            # blank chain everywhere
    molid[nsol + 3 * 10500:] = 20000 + np.arange(30)
    t = device_traj(frames)
    try:
        for com in (False, True):
            got, read = t.rdf_cdf_sum(molid, chain, flags, refidx, com=com, end=12.0)
            want, wread = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, com=com, end=12.0)
            assert read == wread and want.max() > 10
            assert np.array_equal(got, want)
    finally:
        t.close()


def test_rdf_skip_rule_fires_across_runs():
    """Two waters with the same (MolID, chain) and no residue started between them: the reference drops the second one, and a
    third one with the pair again, and counts the next residue with another pair.  Also across the boundary of a work-list chunk
    (more than 4096 candidate atoms between the two) and when the first of the pair is itself dropped."""
    rng = np.random.default_rng(8)
    nsol = 6
    sol = rng.normal(0, 1.0, (nsol, 3))
    near = lambda k: np.array([3.0 + 0.3 * k, 0.0, 0.0]) + rng.normal(0, 0.05, (3, 3))     # a water about 3 A from the solute
    far = lambda: np.array([40.0, 40.0, 40.0]) + rng.normal(0, 3.0, (3, 3))                # beyond end + 3: never starts
    waters, ids = [], []
    # id 7 near, (far ones), id 7 near again -> dropped, id 7 near -> dropped, id 8 near -> counted, id 7 near -> counted again
    for k, (pos, i) in enumerate([(near(0), 7), (far(), 9), (far(), 10), (near(1), 7), (near(2), 7), (near(3), 8), (near(4), 7)]):
        waters.append(pos); ids.append(i)
    # 1500 far waters (4500 atoms: the next near pair straddles a chunk of the work list), then the pair 8 / 8
    for k in range(1500):
        waters.append(far()); ids.append(100 + k)
    waters.append(near(5)); ids.append(3000)
    for k in range(1500):
        waters.append(far()); ids.append(4000 + k)
    waters.append(near(6)); ids.append(3000)          # same pair as the last one started, 4500 atoms earlier -> dropped
    waters.append(near(7)); ids.append(3001)
    coords = np.concatenate([sol] + waters)
    n = coords.shape[0]
    frames = np.stack([coords + rng.normal(0, 0.01, (n, 3)) for _ in range(5)]).astype(np.float32).astype(np.float64)
    molid = np.concatenate([np.ones(nsol), np.repeat(ids, 3)]).astype(np.int32)
    chain = np.zeros(n, dtype=np.int32)
    flags = np.concatenate([np.zeros(nsol), np.full(n - nsol, 3)]).astype(np.uint8)
    refidx = np.arange(nsol)
    t = device_traj(frames)
    try:
        got, read = t.rdf_cdf_sum(molid, chain, flags, refidx, step=0.5, end=10.0)
        want, _ = oracle.rdf_cdf_sum(frames, molid, chain, flags, refidx, step=0.5, end=10.0)
        assert np.array_equal(got, want)
        # 8 near waters, 3 of them dropped by the rule: 5 ranked per frame, minus one (solvation.go:402-404), 5 frames
        assert got[-1] == 5 * (5 - 1)
    finally:
        t.close()
