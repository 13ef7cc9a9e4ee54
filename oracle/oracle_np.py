"""Second, independent restatement of the same path in numpy + LAPACK (scipy gesvd).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED.  scipy's `lapack_driver='gesvd'` is the closest
available stand-in for gonum v0.15.1's Dgesvd (go.mod:10, source absent).  tests/ cross-check
gochem_oracle.c against this file and both against a Horn-quaternion solver (`horn_rotation`),
so three unrelated 3x3 solvers have to agree before the oracle is trusted.

Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


def mass_center(x):
    """chem.MassCenter with unit masses (geometric.go:670-703) -> centred copy, centroid."""
    n = x.shape[0]
    cen = (np.ones((1, n)) @ x) * (1.0 / float(n))
    return x - cen, cen.reshape(3)


def rotator_translator_to_super(test, templa):
    """chem.RotatorTranslatorToSuper (geometric.go:127-208)."""
    test = np.asarray(test, dtype=np.float64)
    templa = np.asarray(templa, dtype=np.float64)
    if test.shape != templa.shape or test.ndim != 2 or test.shape[1] != 3:
        raise ValueError("goChem: Ill-formed matrices")  # :130-132
    n = test.shape[0]
    ctest, distest = mass_center(test)
    ctempla, distempla = mass_center(templa)
    maux = (ctempla.T @ ctest).T  # :149-153  (aux2 = ctempla^T * I)
    u, _s, vt = sla.svd(maux, full_matrices=True, lapack_driver="gesvd")  # :154-165
    v = vt.T
    u, v = -u, -v  # :174-175
    rot = (v @ u.T).T  # :181-182
    if np.linalg.det(rot) < 0:  # :183-188
        rot = (v @ np.diag([1.0, 1.0, -1.0]) @ u.T).T
    jt = np.full((1, n), 1.0 / n)
    translation = (jt @ (ctempla - ctest @ rot)).reshape(3) + distempla  # :191-200
    transformed = ctest @ rot + translation  # :202-204
    return transformed, rot, -distest, translation


def _dispatch(test, templa, indexes):
    """Index dispatch of handy.go:178-201 / geometric.go:261-283."""
    if len(indexes) == 0 or indexes[0] is None or len(indexes[0]) == 0:
        return test, templa
    if len(indexes) == 1:
        if test.shape[0] > len(indexes[0]):
            return test[np.asarray(indexes[0])], templa
        if templa.shape[0] > len(indexes[0]):
            raise RuntimeError("nil dereference in the reference (ctest stays nil)")
        raise ValueError("Indexes don't match molecules")
    return test[np.asarray(indexes[0])], templa[np.asarray(indexes[1])]


def super_(test, templa, *indexes):
    """chem.Super (handy.go:175-214); returns the moved copy of test."""
    test = np.array(test, dtype=np.float64)
    templa = np.asarray(templa, dtype=np.float64)
    ct, cm = _dispatch(test, templa, indexes)
    if ct.shape[0] != cm.shape[0]:
        raise ValueError("chem.Super: Ill formed coordinates for Superposition")
    _, rot, t1, t2 = rotator_translator_to_super(ct, cm)
    return ((test + t1) @ rot) + t2


def mem_rmsd(test, templa, *indexes):
    """chem.MemRMSD (geometric.go:258-288)."""
    ct, cm = _dispatch(np.asarray(test, float), np.asarray(templa, float), indexes)
    if ct.shape[0] != cm.shape[0]:
        raise ValueError("chem.memRMSD: Ill formed matrices")
    return float(np.linalg.norm(ct - cm) / np.sqrt(float(ct.shape[0])))


def per_atom_rmsd(test, templa):
    """chem.MemPerAtomRMSD (geometric.go:294-321)."""
    return np.linalg.norm(np.asarray(test) - np.asarray(templa), axis=1)


def horn_rotation(a_centred, b_centred):
    """Third solver: Horn 1987 unit-quaternion method.  Returns R (right-multiplying row vectors)
    maximising tr(R^T A^T B).  Not from the reference; an independent cross-check only."""
    h = a_centred.T @ b_centred
    sxx, sxy, sxz = h[0]
    syx, syy, syz = h[1]
    szx, szy, szz = h[2]
    k = np.array([
        [sxx + syy + szz, syz - szy, szx - sxz, sxy - syx],
        [syz - szy, sxx - syy - szz, sxy + syx, szx + sxz],
        [szx - sxz, sxy + syx, -sxx + syy - szz, syz + szy],
        [sxy - syx, szx + sxz, syz + szy, -sxx - syy + szz],
    ])
    w, vec = np.linalg.eigh(k)
    q0, q1, q2, q3 = vec[:, np.argmax(w)]
    # rotation acting on column vectors: b ~ Rc a ;  row form R = Rc^T
    rc = np.array([
        [q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
        [2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)],
        [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3],
    ])
    return rc.T


# ---------------------------------------------------------------------------
# LOVO bookkeeping (align/lovo.go), pure Python: small cases only.
# ---------------------------------------------------------------------------

def mobile_limit(limit, tolerance, sorted_rmsd):
    """align.mobileLimit (lovo.go:175-188) with its quirks."""
    n = 0
    guard = 0
    while n < tolerance:
        n = 0
        while n < len(sorted_rmsd):
            if sorted_rmsd[n] >= limit:
                n -= 1
                break
            n += 1
        limit *= 1.1
        guard += 1
        if guard > 100000:
            raise RuntimeError("mobileLimit would not terminate")
    return n, limit


def rmsd_traj(frames, ref, indexes, cpus, begin=0, skip=0):
    """align.RMSDTraj (lovo.go:286-353)."""
    F, N = frames.shape[0], frames.shape[1]
    acc = np.zeros(N)
    b, s = begin // cpus, skip // cpus
    cursor = 0
    chunks = 0
    read = 0
    while True:
        if read < b or read % (s + 1) != 0:
            if cursor + cpus > F:
                break
            cursor += cpus
        if cursor + cpus > F:
            break
        for f in range(cursor, cursor + cpus):
            moved = super_(frames[f], ref, indexes, indexes)
            acc += per_atom_rmsd(moved, ref)
        cursor += cpus
        chunks += 1
        read += 1
    total = float(chunks * cpus)
    with np.errstate(invalid="ignore", divide="ignore"):
        return acc / total, int(total)


def lovo(frames, ref, fullindexes, cpus, less_than_rmsd=1.0, minimum_n=10, max_iter=200):
    """align.LOVO (lovo.go:198-258)."""
    full = list(fullindexes)
    indexesold = list(full)
    n = len(full)
    prevlim = 0.0
    it = 0
    while True:
        rm, total = rmsd_traj(frames, ref, indexesold[:n], cpus)
        it += 1
        vals = [rm[i] for i in range(len(rm)) if i in set(full)]  # rmsdFilter :519-528
        order = sorted(range(len(full)), key=lambda k: vals[k])  # stable, ascending :429-432
        ids = [full[k] for k in order]
        svals = [vals[k] for k in order]
        newlim = 0.0
        if less_than_rmsd > 0:
            n, newlim = mobile_limit(less_than_rmsd, minimum_n, svals)
        same = set(ids[:n]) == set(indexesold[:n]) and len(ids[:n]) == len(indexesold[:n])
        if same and (newlim == less_than_rmsd or abs(newlim) - abs(prevlim) <= 0.01):
            break
        prevlim = newlim
        indexesold = ids
        if it >= max_iter:
            raise RuntimeError("no convergence")
    by_id = sorted(range(len(ids)), key=lambda k: ids[k])
    return dict(N=n, Natoms=np.array(indexesold[:n]), RMSD=np.array([svals[k] for k in by_id]),
                FramesRead=total, Iterations=it)


# ---- solv: molecular RDF, second restatement (numpy, no shared code with the C oracle) ----
def frame_umol_crdf(coord, molid, chain, flags, refidx, com=False, step=0.1, end=10.0):
    """solv.DistRank + FrameUMolCRDF (solvation.go:380-410, 447-541), written from the description of what they compute:
    for every solvent residue (first non-pre-screened atom to the last atom with the same MolID) the minimum distance to the
    solute, kept when <= end; ret[i-1] = max(#{d < i*step} - 1, 0)."""
    coord = np.asarray(coord, dtype=np.float64)
    n = coord.shape[0]
    ref = coord[np.asarray(refidx)]
    own = {(int(molid[i]), int(chain[i])) for i in refidx}
    dists = []
    skip = (-1, 0)
    i = 0
    for i in range(n):
        key = (int(molid[i]), int(chain[i]))
        if flags[i] & 2 and np.linalg.norm(ref[0] - coord[i]) > end + 3:
            continue
        if flags[i] & 1 and key not in own and key != skip:
            j = i + 1
            while j < n and molid[j] == molid[i]:
                j += 1
            test = coord[i:j]
            if com:
                test = (test.sum(axis=0) * (1.0 / (j - i)))[None, :]
            d = np.sqrt(((ref[None, :, :] - test[:, None, :]) ** 2).sum(axis=2)).min()
            d = min(d, 100000.0)
            if d <= end:
                dists.append(d)
            skip = key
    dists = np.sort(np.array(dists))
    out = []
    for k in range(1, int(end / step) + 1):
        c = int(np.searchsorted(dists, k * step, side="left"))
        out.append(float(c - 1 if c > 0 else 0))
    return np.array(out)
