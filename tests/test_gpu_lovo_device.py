"""align.LOVO with the bookkeeping on the device (gcb_lovo_resident, csrc/lovo_ctl.cu) against the same loop driven from the
host: per-atom sums from gcb_peratom_dev_sum, then rmsdFilter / sort.Stable / mobileLimit / sameElementsInt / approxEq as
align/lovo.go:229-252 writes them (the restatement of oracle/oracle_np.py, fed with the GPU's sums).  Every output must be
identical bit for bit: the two loops run the same kernel on the same template."""
import os

import numpy as np
import pytest

import oracle
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H
from oracle import oracle_np

pytestmark = pytest.mark.gpu


def host_driven(t, ref, cand, less, minn, total, max_iter=200):
    full = [int(c) for c in cand]
    indexesold, n, prevlim, it = list(full), len(full), 0.0, 0
    while True:
        sums = t.peratom_dev_sum(ref, np.asarray(indexesold[:n], dtype=np.int32))
        it += 1
        vals = [sums[i] / total for i in full]
        order = sorted(range(len(full)), key=lambda k: vals[k])        # stable
        ids, svals = [full[k] for k in order], [vals[k] for k in order]
        newlim = 0.0
        if less > 0:
            n, newlim = oracle_np.mobile_limit(less, minn, svals)
        same = set(ids[:n]) == set(indexesold[:n])
        if same and (newlim == less or abs(newlim) - abs(prevlim) <= 0.01):
            break
        prevlim, indexesold = newlim, ids
        assert it < max_iter
    return dict(N=n, Iterations=it, Natoms=np.array(indexesold[:n]), ids=np.array(ids), vals=np.array(svals))


CASES = [  # natoms, nframes, candidates, lessThanRMSD, minimumN
    (5000, 4000, "all", 1.0, 10),
    (3000, 3000, "no5", 0.8, 10),
    (2000, 2000, "all", 0.6, 1500),     # mobileLimit has to widen the limit again and again
    (1000, 6000, "even", 0.5, 10),
    (10000, 600, "even", 1.2, 25),     # frames of 10 000 atoms (four CTAs per frame), 5000 candidates
    (12000, 300, "all", 1.0, 10),      # 12 000 candidates: the rank kernel's shared memory beyond 64 KB
    (3000, 1000, "all", 0.0, 10),      # lessThanRMSD <= 0: mobileLimit is skipped, n stays the whole list (lovo.go:238-240)
    (3000, 1000, "no5", 5.0, 10),      # every candidate below the limit: n = the whole list
]


@pytest.mark.parametrize("natoms,nframes,which,less,minn", CASES)
def test_device_loop_equals_host_driven_loop(natoms, nframes, which, less, minn):
    ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
    cand = {"all": np.arange(natoms), "no5": np.array([i for i in range(natoms) if i % 5]), "even": np.arange(0, natoms, 2)}[which].astype(np.int32)
    t = _lib.DeviceTraj(natoms, nframes)
    t.synth_fill(ref, sig, nframes)
    want = host_driven(t, ref, cand, less, minn, float(nframes))
    got = t.lovo_resident(ref, cand, less_than_rmsd=less, minimum_n=minn)
    assert got is not None, "the device loop should take this shape"
    assert got["N"] == want["N"] and got["Iterations"] == want["Iterations"]
    assert np.array_equal(got["Natoms"], want["Natoms"])
    assert np.array_equal(got["ids"], want["ids"])
    assert np.array_equal(got["vals"], want["vals"])           # bit for bit
    # and once more on the same handle: the workspace and the template cache are reused correctly
    again = t.lovo_resident(ref, cand, less_than_rmsd=less, minimum_n=minn)
    assert np.array_equal(again["vals"], got["vals"]) and np.array_equal(again["indexesold"], got["indexesold"])
    # a plain pass afterwards must not see the template the controller left behind
    s1 = t.peratom_dev_sum(ref, cand[: cand.size // 2])
    t2 = _lib.DeviceTraj(natoms, nframes)
    t2.synth_fill(ref, sig, nframes)
    assert np.array_equal(s1, t2.peratom_dev_sum(ref, cand[: cand.size // 2]))
    t.close()
    t2.close()


def test_shapes_the_device_loop_leaves_to_the_host():
    for natoms, cand in ((300, np.arange(300)), (4000, np.arange(0, 4000, 16)), (2000, np.arange(2000)[::-1].copy()), (17000, np.arange(17000))):
        ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
        t = _lib.DeviceTraj(natoms, 64)
        t.synth_fill(ref, sig, 64)
        assert t.lovo_resident(ref, cand.astype(np.int32)) is None
        t.close()


def test_minimum_n_beyond_the_candidates_is_an_error_not_a_hang():
    natoms = 1500
    ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
    t = _lib.DeviceTraj(natoms, 256)
    t.synth_fill(ref, sig, 256)
    with pytest.raises(_lib.GcbError, match="mobileLimit"):
        t.lovo_resident(ref, np.arange(natoms, dtype=np.int32), minimum_n=natoms + 5)
    t.close()


def test_host_mirror_gives_the_same_answer_through_either_loop():
    natoms, nframes = 2500, 4096
    ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
    t = _lib.DeviceTraj(natoms, nframes)
    t.synth_fill(ref, sig, nframes)
    dev = H.LOVOResident(t, ref, cpus=8, less_than_rmsd=0.9, minimum_n=10)
    os.environ["GCH_LOVO_HOST"] = "1"
    try:
        host = H.LOVOResident(t, ref, cpus=8, less_than_rmsd=0.9, minimum_n=10)
    finally:
        del os.environ["GCH_LOVO_HOST"]
    assert dev["N"] == host["N"] and dev["Iterations"] == host["Iterations"]
    assert np.array_equal(dev["Natoms"], host["Natoms"]) and np.array_equal(dev["RMSD"], host["RMSD"])
    frames = t.download()
    want = oracle.lovo(frames, ref, np.arange(natoms), cpus=8, less_than_rmsd=0.9, minimum_n=10, nthreads=16)
    assert dev["N"] == want["N"] and dev["Iterations"] == want["Iterations"] and list(dev["Natoms"]) == list(want["Natoms"])
    assert np.abs(dev["RMSD"] - want["RMSD"]).max() < 1e-10
    t.close()


def test_two_handles_from_two_threads():
    """The device loop keeps its workspace, staging words and template in the handle: two handles driven from two threads at the
    same time give what they give one after the other."""
    import threading

    cases = [(3000, 2000, 0.8), (5000, 1500, 1.0)]
    handles, refs, serial = [], [], []
    for natoms, nframes, less in cases:
        ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
        t = _lib.DeviceTraj(natoms, nframes)
        t.synth_fill(ref, sig, nframes)
        handles.append(t)
        refs.append(ref)
        serial.append(t.lovo_resident(ref, np.arange(natoms, dtype=np.int32), less_than_rmsd=less))
    out = [None, None]

    def work(k):
        natoms, _, less = cases[k]
        for _ in range(5):
            out[k] = handles[k].lovo_resident(refs[k], np.arange(natoms, dtype=np.int32), less_than_rmsd=less)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for k in range(2):
        assert out[k]["N"] == serial[k]["N"] and out[k]["Iterations"] == serial[k]["Iterations"]
        assert np.array_equal(out[k]["indexesold"], serial[k]["indexesold"]) and np.array_equal(out[k]["vals"], serial[k]["vals"])
        handles[k].close()
