"""align.LOVO at the atom count of BASELINE.json configs[2] (5,000 atoms; 20,000 of the 50,000 frames to keep the
oracle's CPU time to seconds) on a device-generated trajectory: N, the ordered Natoms list, the iteration count
and the per-atom means must equal the oracle's on the exact bytes the GPU holds."""
import numpy as np
import pytest

import oracle
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

pytestmark = pytest.mark.gpu


def test_lovo_config3_shape():
    natoms, nframes, cpus = 5000, 20000, 8
    ref = synth.make_reference(natoms)
    sig = synth.atom_sigmas(natoms)
    t = _lib.DeviceTraj(natoms, nframes)
    t.synth_fill(ref, sig, nframes)
    got = H.LOVOResident(t, ref, cpus=cpus, less_than_rmsd=1.0, minimum_n=10)
    frames = t.download()
    want = oracle.lovo(frames, ref, np.arange(natoms), cpus=cpus, less_than_rmsd=1.0, minimum_n=10, nthreads=16)
    assert got["N"] == want["N"] and got["Iterations"] == want["Iterations"]
    assert list(got["Natoms"]) == list(want["Natoms"])
    assert np.abs(got["RMSD"] - want["RMSD"]).max() < 1e-10
    assert 100 < got["N"] < natoms  # the synthetic noise model has rigid and mobile atoms
    t.close()
