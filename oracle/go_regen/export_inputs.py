"""Inputs for oracle/go_regen/main.go: every superposition case of tests/golden as a flat little-endian file.
    python oracle/go_regen/export_inputs.py     ->  tests/golden/go_reference/<case>.in.bin
Test infrastructure only (see oracle/__init__.py)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(GOLDEN, "go_reference")
CPUS = 3   # chunk size of align.RMSDTraj in the Go run; frames beyond the last full chunk are dropped (lovo.go:323-326)


def export(path: str) -> str:
    z = np.load(path)
    ref = z["ref"].astype("<f8")
    frames = z["frames"].astype("<f8")
    idx = z["idx"].astype("<i8") if "idx" in z.files else np.zeros(0, "<i8")
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, os.path.basename(path)[:-4] + ".in.bin")
    with open(out, "wb") as f:
        np.array([ref.shape[0], frames.shape[0], idx.size, CPUS], "<i8").tofile(f)
        ref.tofile(f)
        frames.tofile(f)
        idx.tofile(f)
    return out


def read_outputs(path: str, natoms: int, nframes: int) -> dict:
    """<case>.go.bin as written by main.go."""
    with open(path, "rb") as f:
        def take(n, dt="<f8"):
            return np.fromfile(f, dt, n)
        out = {"rmsd": take(nframes), "rot": take(nframes * 9).reshape(nframes, 3, 3), "trans1": take(nframes * 3).reshape(nframes, 3),
               "trans2": take(nframes * 3).reshape(nframes, 3), "moved": take(nframes * natoms * 3).reshape(nframes, natoms, 3)}
        out["frames_read"] = int(take(1, "<i8")[0])
        out["rmsdtraj"] = take(natoms)
    return out


def export_lovo(path: str) -> str:
    """The align.LOVO case: same layout, no index list, the chunk size of the golden run."""
    z = np.load(path)
    ref, frames = z["ref"].astype("<f8"), z["frames"].astype("<f8")
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, os.path.basename(path)[:-4] + ".lovo.in.bin")
    with open(out, "wb") as f:
        np.array([ref.shape[0], frames.shape[0], 0, int(z["cpus"])], "<i8").tofile(f)
        ref.tofile(f)
        frames.tofile(f)
        if "begin_skip" in z.files:   # (begin, skip) pairs of the align.RMSDTraj runs
            combos = z["begin_skip"].astype("<i8")
            np.array([combos.shape[0]], "<i8").tofile(f)
            combos.tofile(f)
    return out


def export_rdf(path: str) -> str:
    """The solv.FrameUMolCRDF case: topology as three integer arrays, the solute, float64 frames."""
    z = np.load(path)
    frames = z["frames"].astype("<f8")
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, os.path.basename(path)[:-4] + ".rdf.in.bin")
    with open(out, "wb") as f:
        np.array([frames.shape[1], frames.shape[0], z["refidx"].size], "<i8").tofile(f)
        np.array([float(z["step"]), float(z["end"])], "<f8").tofile(f)
        for k in ("molid", "chain", "flags", "refidx"):
            z[k].astype("<i8").tofile(f)
        frames.tofile(f)
    return out


def read_moments(path: str, nframes: int) -> dict:
    a = np.fromfile(path, "<f8").reshape(nframes, 12)
    return {"center": a[:, :3], "moment": a[:, 3:].reshape(nframes, 3, 3)}


def read_rmsdtraj(path: str, natoms: int, ncombos: int) -> list:
    out = []
    with open(path, "rb") as f:
        for _ in range(ncombos):
            read = int(np.fromfile(f, "<i8", 1)[0])
            out.append((read, np.fromfile(f, "<f8", natoms)))
    return out


def read_rdf(path: str, nframes: int, nsteps: int) -> dict:
    a = np.fromfile(path, "<f8").reshape(2, nframes, nsteps)
    return {"com0": a[0], "com1": a[1]}


def read_lovo_outputs(path: str) -> dict:
    """<case>.lovo.go.bin as written by main.go."""
    with open(path, "rb") as f:
        n, frames_read, iterations, nnat, nrmsd = (int(v) for v in np.fromfile(f, "<i8", 5))
        return {"N": n, "FramesRead": frames_read, "Iterations": iterations, "Natoms": np.fromfile(f, "<i8", nnat),
                "RMSD": np.fromfile(f, "<f8", nrmsd)}


if __name__ == "__main__":
    for p in sorted(glob.glob(os.path.join(GOLDEN, "super_*.npz"))):
        print(export(p))
    for p in sorted(glob.glob(os.path.join(GOLDEN, "lovo_*.npz"))):
        print(export_lovo(p))
    for p in sorted(glob.glob(os.path.join(GOLDEN, "rdf_*.npz"))):
        print(export_rdf(p))
