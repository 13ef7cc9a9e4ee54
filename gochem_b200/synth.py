"""Seeded synthetic trajectories (SURVEY.md §8d / BASELINE.md §4).

Host-side numpy generator for the small parity cases.  The large bench configurations are
generated on the device by `gcb_synth_fill` (csrc/synth.cu) from the same recipe with a
counter-based RNG; the two generators are NOT bit-identical, so full-size parity checks
download the sampled frames and hand those exact bytes to the oracle.

Recipe: reference = N points ~ N(0, 12^2) per axis (seed 12345); per-atom noise sigma_a
log-uniform in [0.2, 3] A (seed 12346); frame f = (reference + noise_f) rotated by a random
unit quaternion and translated uniformly in [-50, 50]^3 A (seed 12347 + f).
"""
from __future__ import annotations

import numpy as np

REF_SEED, SIGMA_SEED, FRAME_SEED = 12345, 12346, 12347
REF_SIGMA = 12.0


def make_reference(natoms: int, seed: int = REF_SEED) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((natoms, 3)) * REF_SIGMA


def atom_sigmas(natoms: int, seed: int = SIGMA_SEED, lo: float = 0.2, hi: float = 3.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.exp(rng.uniform(np.log(lo), np.log(hi), natoms))


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """Proper rotation (det +1) from a unit quaternion; rows act on row vectors from the right."""
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def make_frames(ref: np.ndarray, sigmas: np.ndarray, nframes: int, seed0: int = FRAME_SEED,
                frame0_is_ref: bool = False, through_f32: bool = False, first_frame: int = 0) -> np.ndarray:
    """F x N x 3 float64 (row-major AoS, the v3.Matrix layout, v3/gonum.go:54-58)."""
    n = ref.shape[0]
    out = np.empty((nframes, n, 3), dtype=np.float64)
    for k in range(nframes):
        f = first_frame + k
        if f == 0 and frame0_is_ref:
            out[k] = ref
            continue
        rng = np.random.Generator(np.random.PCG64(seed0 + f))
        noisy = ref + rng.standard_normal((n, 3)) * sigmas[:, None]
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        t = rng.uniform(-50.0, 50.0, 3)
        out[k] = noisy @ quat_to_rot(q) + t
    if through_f32:
        out = out.astype(np.float32).astype(np.float64)
    return out


def make_case(natoms: int, nframes: int, frame0_is_ref: bool = False, through_f32: bool = False):
    ref = make_reference(natoms)
    if through_f32:
        ref = ref.astype(np.float32).astype(np.float64)
    sig = atom_sigmas(natoms)
    return ref, make_frames(ref, sig, nframes, frame0_is_ref=frame0_is_ref, through_f32=through_f32)
