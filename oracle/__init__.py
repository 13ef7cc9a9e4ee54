"""ctypes binding of the CPU oracle (oracle/gochem_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see the C file's header).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; nothing under gochem_b200/ does.

Each wrapper names the reference function it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

OK, ERR_SHAPE, ERR_INDEXES, ERR_NILDEREF, ERR_SVD, ERR_ALLOC, ERR_NOTERM = 0, -1, -2, -3, -4, -5, -6


def build(force: bool = False) -> str:
    """Compile liborc.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "gochem_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rmsd_naive.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a, copy=False):
    a = np.array(a, dtype=np.float64, order="C", copy=True) if copy else np.ascontiguousarray(a, dtype=np.float64)
    return a


def _idx(a):
    if a is None:
        return None, 0
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, int(a.size)


class OracleError(Exception):
    def __init__(self, code):
        super().__init__({ERR_SHAPE: "ill-formed matrices", ERR_INDEXES: "indexes don't match molecules",
                          ERR_NILDEREF: "reference would nil-dereference ctest", ERR_SVD: "mat.SVD failed",
                          ERR_ALLOC: "alloc", ERR_NOTERM: "would not terminate"}.get(code, str(code)))
        self.code = code


def svd3(h):
    """Stand-in for gonum mat.SVD on the 3x3 correlation (geometric.go:154-165)."""
    h = _f64(h).reshape(3, 3)
    u, s, v = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    e = lib().orc_svd3(_d(h), _d(u), _d(s), _d(v))
    if e:
        raise OracleError(e)
    return u, s, v


def rotator_translator_to_super(test, templa, as_written=False):
    """chem.RotatorTranslatorToSuper (geometric.go:127-208) -> transformed, R, trans1, trans2."""
    test, templa = _f64(test), _f64(templa)
    if test.shape != templa.shape or test.ndim != 2 or test.shape[1] != 3:
        raise OracleError(ERR_SHAPE)
    n = test.shape[0]
    tr, r, t1, t2 = np.zeros((n, 3)), np.zeros((3, 3)), np.zeros(3), np.zeros(3)
    e = lib().orc_rotator_translator(_d(test), _d(templa), n, _d(tr), _d(r), _d(t1), _d(t2), int(as_written))
    if e:
        raise OracleError(e)
    return tr, r, t1, t2


def super_(test, templa, *indexes, as_written=False, details=False):
    """chem.Super (handy.go:175-214).  Returns the superposed copy of test (the reference mutates
    test in place; the copy keeps the caller's array intact)."""
    out = _f64(test, copy=True)
    templa = _f64(templa)
    i0, n0 = _idx(indexes[0]) if len(indexes) > 0 else (None, 0)
    i1, n1 = _idx(indexes[1]) if len(indexes) > 1 else (None, 0)
    r, t1, t2 = np.zeros((3, 3)), np.zeros(3), np.zeros(3)
    e = lib().orc_super(_d(out), out.shape[0], _d(templa), templa.shape[0],
                        _i(i0) if i0 is not None else None, n0, _i(i1) if i1 is not None else None, n1,
                        len(indexes), int(as_written), _d(r), _d(t1), _d(t2))
    if e:
        raise OracleError(e)
    return (out, r, t1, t2) if details else out


def mem_rmsd(test, templa, *indexes):
    """chem.RMSD / chem.MemRMSD (geometric.go:232-288): no fit; Frobenius / sqrt(n)."""
    test, templa = _f64(test), _f64(templa)
    i0, n0 = _idx(indexes[0]) if len(indexes) > 0 else (None, 0)
    i1, n1 = _idx(indexes[1]) if len(indexes) > 1 else (None, 0)
    out = C.c_double(0)
    e = lib().orc_mem_rmsd(_d(test), test.shape[0], _d(templa), templa.shape[0],
                           _i(i0) if i0 is not None else None, n0, _i(i1) if i1 is not None else None, n1,
                           len(indexes), C.byref(out))
    if e:
        raise OracleError(e)
    return out.value


rmsd = mem_rmsd


def per_atom_rmsd(test, templa):
    """chem.MemPerAtomRMSD without index lists (geometric.go:294-321): d_i = |test_i - templa_i|."""
    test, templa = _f64(test), _f64(templa)
    out = np.zeros(test.shape[0])
    lib().orc_per_atom_rmsd(_d(test), _d(templa), test.shape[0], _d(out))
    return out


def rmsd_naive(test, templa):
    """chem.rMSD (geometric.go:326-364)."""
    test, templa = _f64(test), _f64(templa)
    return lib().orc_rmsd_naive(_d(test), _d(templa), test.shape[0])


def rmsd_traj(frames, ref, indexes, cpus, begin=0, skip=0, nthreads=1, as_written=False, want_aligned=False):
    """align.RMSDTraj (align/lovo.go:286-353) on an in-memory F x N x 3 trajectory."""
    frames, ref = _f64(frames), _f64(ref)
    F, N = frames.shape[0], frames.shape[1]
    idx, nidx = _idx(indexes)
    out = np.zeros(N)
    fr = C.c_long(0)
    aligned = np.zeros((F - F % cpus, N, 3)) if want_aligned else None
    e = lib().orc_rmsd_traj(_d(frames), C.c_long(F), N, _d(ref), _i(idx), nidx, cpus, begin, skip, nthreads,
                            int(as_written), _d(out), C.byref(fr), _d(aligned) if want_aligned else None)
    if e:
        raise OracleError(e)
    if want_aligned:
        return out, fr.value, aligned[: fr.value]
    return out, fr.value


def super_rmsd_traj(frames, ref, indexes=None, write_back=False, as_written=False, nthreads=1, details=True):
    """Per-frame chem.Super + chem.RMSD (the user loop of gochem_test.go:471-473)."""
    frames = _f64(frames, copy=not write_back)
    ref = _f64(ref)
    F, N = frames.shape[0], frames.shape[1]
    idx, nidx = _idx(indexes)
    rm = np.zeros(F)
    rot, t1, t2 = (np.zeros((F, 3, 3)), np.zeros((F, 3)), np.zeros((F, 3))) if details else (None, None, None)
    e = lib().orc_super_rmsd_traj(_d(frames), C.c_long(F), N, _d(ref), _i(idx) if idx is not None else None, nidx,
                                  int(write_back), int(as_written), nthreads, _d(rm),
                                  _d(rot) if details else None, _d(t1) if details else None,
                                  _d(t2) if details else None)
    if e:
        raise OracleError(e)
    return (rm, rot, t1, t2) if details else rm


def mobile_limit(limit, tolerance, sorted_rmsd):
    """align.mobileLimit (lovo.go:175-188)."""
    v = _f64(sorted_rmsd)
    n, lim = C.c_int(0), C.c_double(0)
    e = lib().orc_mobile_limit(C.c_double(limit), tolerance, _d(v), v.size, C.byref(n), C.byref(lim))
    if e:
        raise OracleError(e)
    return n.value, lim.value


def stable_sort_pairs(ids, vals, by_rmsd=True):
    ids = np.array(ids, dtype=np.int32)
    vals = np.array(vals, dtype=np.float64)
    lib().orc_stable_sort_pairs(_i(ids), _d(vals), ids.size, int(by_rmsd))
    return ids, vals


def lovo(frames, ref, fullindexes, cpus, less_than_rmsd=1.0, minimum_n=10, begin=0, skip=0, nthreads=1,
         max_iter=200):
    """align.LOVO (lovo.go:198-258) on an in-memory trajectory.
    Returns dict(N, Natoms, RMSD, FramesRead, Iterations)."""
    frames, ref = _f64(frames), _f64(ref)
    F, N = frames.shape[0], frames.shape[1]
    full, nfull = _idx(fullindexes)
    n, it, fr = C.c_int(0), C.c_int(0), C.c_long(0)
    natoms = np.zeros(max(nfull, 1), dtype=np.int32)
    rm = np.zeros(max(nfull, 1))
    e = lib().orc_lovo(_d(frames), C.c_long(F), N, _d(ref), _i(full), nfull, cpus, begin, skip,
                       C.c_double(less_than_rmsd), minimum_n, nthreads, max_iter, C.byref(n), _i(natoms), _d(rm),
                       C.byref(fr), C.byref(it), None)
    if e:
        raise OracleError(e)
    return dict(N=n.value, Natoms=natoms[: n.value].copy(), RMSD=rm[:nfull].copy(), FramesRead=fr.value,
                Iterations=it.value)


def center_and_moment(a, mass=None):
    """chem.CenterOfMass (geometric.go:609-631) and chem.MomentTensor (geometric.go:705-733)."""
    a = _f64(a)
    m = None if mass is None else _f64(mass)
    cen, mom = np.zeros(3), np.zeros((3, 3))
    e = lib().orc_center_and_moment(_d(a), a.shape[0], _d(m) if m is not None else None, _d(cen), _d(mom))
    if e:
        raise OracleError(e)
    return cen, mom


# ---- solv: molecular RDF (solv/solvation.go) ----
def _topo(molid, chain, flags):
    return (np.ascontiguousarray(molid, dtype=np.int32), np.ascontiguousarray(chain, dtype=np.int32),
            np.ascontiguousarray(flags, dtype=np.uint8))


def dist_rank(coord, molid, chain, flags, refidx, com=False, end=10.0):
    """solv.DistRank (solvation.go:447-541) -> (distances, molids) in scan order."""
    coord = _f64(coord)
    m, c, f = _topo(molid, chain, flags)
    r, nr = _idx(refidx)
    n = coord.shape[0]
    d, ids = np.empty(max(n, 1)), np.empty(max(n, 1), dtype=np.int32)
    cnt = lib().orc_dist_rank(_d(coord), n, _i(m), _i(c), f.ctypes.data_as(C.POINTER(C.c_ubyte)), _i(r), nr, int(com),
                              C.c_double(end), _d(d), _i(ids))
    return d[:cnt].copy(), ids[:cnt].copy()


def frame_umol_crdf(coord, molid, chain, flags, refidx, com=False, step=0.1, end=10.0):
    """solv.FrameUMolCRDF (solvation.go:380-410)."""
    coord = _f64(coord)
    m, c, f = _topo(molid, chain, flags)
    r, nr = _idx(refidx)
    ret = np.empty(int(end / step) + 1)
    n = lib().orc_frame_umol_crdf(_d(coord), coord.shape[0], _i(m), _i(c), f.ctypes.data_as(C.POINTER(C.c_ubyte)), _i(r), nr,
                                  int(com), C.c_double(step), C.c_double(end), _d(ret))
    return ret[:n].copy()


def rdf_cdf_sum(frames, molid, chain, flags, refidx, com=False, step=0.1, end=10.0, cpus=1):
    """The frame loop of solv.ConcMolRDF (solvation.go:166-214) -> (summed cumulative counts, frames read)."""
    frames = _f64(frames)
    m, c, f = _topo(molid, chain, flags)
    r, nr = _idx(refidx)
    out = np.empty(int(end / step) + 1)
    fn = lib().orc_rdf_cdf_sum
    fn.restype = C.c_long
    read = fn(_d(frames), C.c_long(frames.shape[0]), frames.shape[1], _i(m), _i(c), f.ctypes.data_as(C.POINTER(C.c_ubyte)), _i(r), nr,
              int(com), C.c_double(step), C.c_double(end), int(cpus), _d(out))
    return out[:int(end / step)].copy(), int(read)


def mdf_from_cdf(cdf_sum, framesread, A, B, step):
    """solv.MDFFromCDF (solvation.go:317-352) -> (normalised density, average cumulative number)."""
    ret = _f64(cdf_sum, copy=True)
    ret2 = np.empty_like(ret)
    rc = lib().orc_mdf_from_cdf(_d(ret), ret.size, int(framesread), C.c_double(A), C.c_double(B), C.c_double(step), _d(ret2))
    if rc:
        raise OracleError(rc)
    return ret, ret2
