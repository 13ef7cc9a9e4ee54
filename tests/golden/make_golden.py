"""Regenerates tests/golden/*.npz.

The Go reference cannot run here (no Go toolchain), so these vectors come from oracle/oracle_np.py
(numpy + LAPACK gesvd, the closest stand-in for gonum's Dgesvd) on seeded synthetic inputs.  They pin
the oracle and the CUDA path against accidental change; they are NOT outputs of the reference itself
(parity unpinned, DESIGN.md "Oracle").   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gochem_b200 import synth  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def super_case(name, natoms, nframes, idx=None, frame0_is_ref=True):
    ref, frames = synth.make_case(natoms, nframes, frame0_is_ref=frame0_is_ref, through_f32=True)
    rm, rot, t1, t2 = [], [], [], []
    moved = []
    for f in range(nframes):
        if idx is None:
            _, r, a, b = onp.rotator_translator_to_super(frames[f], ref)
            m = onp.super_(frames[f], ref)
            rm.append(onp.mem_rmsd(m, ref))
        else:
            _, r, a, b = onp.rotator_translator_to_super(frames[f][idx], ref[idx])
            m = onp.super_(frames[f], ref, idx, idx)
            rm.append(onp.mem_rmsd(m, ref, idx, idx))
        rot.append(r); t1.append(a); t2.append(b); moved.append(m)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ref=ref.astype(np.float32), frames=frames.astype(np.float32),
                        idx=np.array([] if idx is None else idx, dtype=np.int32), rmsd=np.array(rm), rot=np.array(rot),
                        trans1=np.array(t1), trans2=np.array(t2), moved_first=moved[1], moved_last=moved[-1])


def edge_case(name, kind, natoms, nframes):
    """The reflection / rank-deficient branch of geometric.go:183-188 (tests/test_gpu_edge_cases.py makes the inputs):
    mirror images of the template, planar sets, a near-degenerate pair of singular values."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_edge_cases import make_edge_case

    ref, frames = make_edge_case(kind, natoms, nframes)
    ref = ref.astype(np.float32).astype(np.float64)
    frames = frames.astype(np.float32).astype(np.float64)
    rm, rot, t1, t2, moved = [], [], [], [], []
    for f in range(nframes):
        _, r, a, b = onp.rotator_translator_to_super(frames[f], ref)
        m = onp.super_(frames[f], ref)
        rm.append(onp.mem_rmsd(m, ref)); rot.append(r); t1.append(a); t2.append(b); moved.append(m)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ref=ref.astype(np.float32), frames=frames.astype(np.float32),
                        idx=np.array([], dtype=np.int32), rmsd=np.array(rm), rot=np.array(rot), trans1=np.array(t1), trans2=np.array(t2),
                        moved_first=moved[1], moved_last=moved[-1])


def lovo_case(name, natoms, nframes, cpus):
    ref, frames = synth.make_case(natoms, nframes, through_f32=True)
    res = onp.lovo(frames, ref, list(range(natoms)), cpus, less_than_rmsd=1.0, minimum_n=10)
    rm, total = onp.rmsd_traj(frames, ref, list(range(natoms)), cpus)
    # align.RMSDTraj with Begin / Skip (the chunk bookkeeping of lovo.go:306-326): (begin, skip) pairs the Go program repeats
    combos = np.array([[0, 0], [4, 0], [0, 4], [8, 8], [5, 3]], dtype=np.int32)
    bs_rmsd, bs_frames = [], []
    for b, k in combos:
        r, t = onp.rmsd_traj(frames, ref, list(range(0, natoms, 2)), cpus, begin=int(b), skip=int(k))
        bs_rmsd.append(r); bs_frames.append(t)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ref=ref.astype(np.float32), frames=frames.astype(np.float32),
                        cpus=cpus, N=res["N"], Natoms=res["Natoms"], RMSD=res["RMSD"], FramesRead=res["FramesRead"],
                        Iterations=res["Iterations"], first_pass_rmsd=rm, first_pass_frames=total,
                        begin_skip=combos, begin_skip_rmsd=np.array(bs_rmsd), begin_skip_frames=np.array(bs_frames))


def rdf_case(name, nsolute=40, nwater=220, nions=12, nframes=10, seed=21, box=17.0):
    """solv.ConcMolRDF's frame loop on a small solvated system, from the numpy restatement (oracle_np.frame_umol_crdf)."""
    rng = np.random.default_rng(seed)
    sol = rng.normal(0, 4, (nsolute, 3))
    wo = rng.uniform(-box, box, (nwater, 3))
    wat = (wo[:, None, :] + np.concatenate([np.zeros((nwater, 1, 3)), rng.normal(0, 0.6, (nwater, 2, 3))], axis=1)).reshape(-1, 3)
    base = np.concatenate([sol, wat, rng.uniform(-box, box, (nions, 3))])
    frames = (base[None] + rng.normal(0, 0.5, (nframes,) + base.shape)).astype(np.float32)
    molid = np.concatenate([np.repeat(np.arange(1, 5), nsolute // 4), 5 + np.repeat(np.arange(nwater), 3), 5 + nwater + np.arange(nions)]).astype(np.int32)
    chain = np.concatenate([np.ones(nsolute), np.full(3 * nwater + nions, 2)]).astype(np.int32)
    flags = np.concatenate([np.zeros(nsolute), np.full(3 * nwater, 3), np.full(nions, 1)]).astype(np.uint8)
    refidx = np.arange(nsolute, dtype=np.int32)
    out = {}
    for com in (0, 1):
        per = np.array([onp.frame_umol_crdf(frames[f].astype(np.float64), molid, chain, flags, refidx, com=bool(com), step=0.1, end=10.0)
                        for f in range(nframes)])
        out["cdf_frames_com%d" % com] = per
    np.savez_compressed(os.path.join(OUT, name + ".npz"), frames=frames, molid=molid, chain=chain, flags=flags, refidx=refidx,
                        step=0.1, end=10.0, **out)


if __name__ == "__main__":
    rdf_case("rdf_small")
    super_case("super_all_n50", 50, 6)
    super_case("super_subset_n203", 203, 5, idx=list(range(3, 203, 4)))
    super_case("super_cfg1_sample_n1500", 1500, 6)
    for kind in ("mirror", "planar", "neardeg"):
        edge_case("super_edge_" + kind + "_n120", kind, 120, 8)
    lovo_case("lovo_n150_f42", 150, 42, 4)
    print("written", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
