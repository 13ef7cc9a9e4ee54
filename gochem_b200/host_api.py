"""ctypes binding of the C shim (include/gochem_host.h) over the C++ host-side mirror of the reference's Go API.
Test / example plumbing: the functions read like the reference's (chem.Super, chem.RMSD, align.LOVO ...)."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgochem_host.so")
HEADER = os.path.join(HERE, "..", "include", "gochem_host.h")


class GoChemError(RuntimeError):
    def __init__(self, code, msg, last_frame=False):
        super().__init__(msg)
        self.code, self.msg, self.last_frame = code, msg, last_frame


_lib = None


def declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(gch_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GoChemError(-7, f"{LIB_PATH} missing: run __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        lib.gch_last_error.restype = C.c_size_t
        lib.gch_selected_frames.restype = C.c_long
        lib.gch_lovo_state_new.restype = C.c_void_p
        lib.gch_shard_range.restype = None
        lib.gch_lovo_state_free.restype = None
        _lib = lib
    return _lib


def _err(rc):
    buf = C.create_string_buffer(4096)
    load().gch_last_error(buf, 4096)
    return GoChemError(rc, buf.value.decode(errors="replace"), bool(load().gch_last_error_is_last_frame()))


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _lists(indexes):
    out = []
    for ix in indexes[:2]:
        out.append(None if ix is None else np.ascontiguousarray(ix, dtype=np.int32))
    while len(out) < 2:
        out.append(None)
    n = [0 if a is None else a.size for a in out]
    return out[0], n[0], out[1], n[1], len(indexes)


def Super(test, templa, *indexes):
    """chem.Super: returns the superposed copy (the Go function mutates test and returns it)."""
    t = np.array(test, dtype=np.float64, order="C")
    m = np.ascontiguousarray(templa, dtype=np.float64)
    i0, n0, i1, n1, nl = _lists(indexes)
    rc = load().gch_super(_d(t), t.shape[0], _d(m), m.shape[0], _i(i0), n0, _i(i1), n1, nl)
    if rc:
        raise _err(rc)
    return t


def RMSD(test, templa, *indexes):
    t = np.ascontiguousarray(test, dtype=np.float64)
    m = np.ascontiguousarray(templa, dtype=np.float64)
    i0, n0, i1, n1, nl = _lists(indexes)
    out = C.c_double(0)
    rc = load().gch_rmsd(_d(t), t.shape[0], _d(m), m.shape[0], _i(i0), n0, _i(i1), n1, nl, C.byref(out))
    if rc:
        raise _err(rc)
    return out.value


def MemPerAtomRMSD(test, templa, *indexes):
    t = np.ascontiguousarray(test, dtype=np.float64)
    m = np.ascontiguousarray(templa, dtype=np.float64)
    i0, n0, i1, n1, nl = _lists(indexes)
    out = np.zeros(n0 if (nl and n0) else t.shape[0])
    rc = load().gch_mem_per_atom_rmsd(_d(t), t.shape[0], _d(m), m.shape[0], _i(i0), n0, _i(i1), n1, nl, _d(out))
    if rc:
        raise _err(rc)
    return out


def RotatorTranslatorToSuper(test, templa):
    t = np.ascontiguousarray(test, dtype=np.float64)
    m = np.ascontiguousarray(templa, dtype=np.float64)
    n = t.shape[0]
    tr, r, a, b = np.zeros((n, 3)), np.zeros((3, 3)), np.zeros(3), np.zeros(3)
    # shape mismatches are the library's to report (geometric.go:130-132), not this wrapper's
    rc = load().gch_rotator_translator_to_super(_d(t), n, _d(m), m.shape[0], _d(tr), _d(r), _d(a), _d(b))
    if rc:
        raise _err(rc)
    return tr, r, a, b


def CenterAndMoment(a, mass=None):
    """chem.CenterOfMass + chem.MomentTensor of one structure."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    m = None if mass is None else np.ascontiguousarray(mass, dtype=np.float64)
    cen, mom = np.zeros(3), np.zeros((3, 3))
    rc = load().gch_center_and_moment(_d(a), a.shape[0], _d(m), _d(cen), _d(mom))
    if rc:
        raise _err(rc)
    return cen, mom


def mobile_limit(limit, tolerance, sorted_rmsd):
    v = np.ascontiguousarray(sorted_rmsd, dtype=np.float64)
    n, lim = C.c_int(0), C.c_double(0)
    rc = load().gch_mobile_limit(C.c_double(limit), tolerance, _d(v), v.size, C.byref(n), C.byref(lim))
    if rc:
        raise _err(rc)
    return n.value, lim.value


def selected_frames(F, cpus, begin=0, skip=0):
    n = load().gch_selected_frames(C.c_long(F), cpus, begin, skip, None, C.c_long(0))
    out = np.zeros(max(n, 1), dtype=np.int64)
    load().gch_selected_frames(C.c_long(F), cpus, begin, skip, out.ctypes.data_as(C.POINTER(C.c_long)), C.c_long(n))
    return out[:n]


def shard_range(nsel, rank, nranks):
    f, c = C.c_long(0), C.c_long(0)
    load().gch_shard_range(C.c_long(nsel), rank, nranks, C.byref(f), C.byref(c))
    return f.value, c.value


class LovoState:
    def __init__(self, fullindexes, less_than_rmsd=1.0, minimum_n=10):
        full = np.ascontiguousarray(fullindexes, dtype=np.int32)
        self.nfull = full.size
        self.h = C.c_void_p(load().gch_lovo_state_new(_i(full), full.size, C.c_double(less_than_rmsd), minimum_n))

    def fit_indexes(self):
        out = np.zeros(max(self.nfull, 1), dtype=np.int32)
        n = load().gch_lovo_state_fit_indexes(self.h, _i(out), out.size)
        return out[:n].copy()

    def step(self, rmsd):
        r = np.ascontiguousarray(rmsd, dtype=np.float64)
        conv = C.c_int(0)
        rc = load().gch_lovo_state_step(self.h, _d(r), r.size, C.byref(conv))
        if rc:
            raise _err(rc)
        return bool(conv.value)

    def disagreement(self):
        return load().gch_lovo_state_disagreement(self.h)

    def finish(self, natoms, frames_read):
        N, it = C.c_int(0), C.c_int(0)
        nat = np.zeros(max(self.nfull, 1), dtype=np.int32)
        rm = np.zeros(max(self.nfull, 1))
        load().gch_lovo_state_finish(self.h, natoms, frames_read, C.byref(N), _i(nat), _d(rm), C.byref(it))
        return dict(N=N.value, Natoms=nat[: N.value].copy(), RMSD=rm[: self.nfull].copy(), FramesRead=frames_read,
                    Iterations=it.value)

    def __del__(self):
        try:
            load().gch_lovo_state_free(self.h)
        except Exception:
            pass


def dcd_write(path, frames):
    f = np.ascontiguousarray(frames, dtype=np.float64)
    rc = load().gch_dcd_write(path.encode(), _d(f), C.c_long(f.shape[0]), f.shape[1])
    if rc:
        raise _err(rc)


def dcd_read(path, cap_frames, natoms, chunk=0):
    out = np.zeros((cap_frames, natoms, 3))
    nf, na = C.c_long(0), C.c_int(0)
    rc = load().gch_dcd_read(path.encode(), _d(out), C.c_long(cap_frames), C.byref(nf), C.byref(na), chunk)
    if rc:
        raise _err(rc)
    return out[: min(nf.value, cap_frames)], nf.value, na.value


def RMSDTraj(dcd_path, ref, indexes, cpus, begin=0, skip=0, write_traj="", dev=0, rank=0, nranks=1, comm=None):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    idx = np.ascontiguousarray(indexes, dtype=np.int32)
    out = np.zeros(ref.shape[0])
    frames = C.c_int(0)
    rc = load().gch_rmsd_traj(dcd_path.encode(), _d(ref), ref.shape[0], _i(idx), idx.size, cpus, begin, skip,
                              write_traj.encode(), dev, rank, nranks, comm, _d(out), C.byref(frames))
    if rc:
        raise _err(rc)
    return out, frames.value


def LOVO(dcd_path, ref, cpus, less_than_rmsd=1.0, minimum_n=10, names=None, chains=None, molids=None, atom_name="CA",
         write_traj="", dev=0, rank=0, nranks=1, comm=None):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    n = ref.shape[0]
    N, frames, it, nr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    nat = np.zeros(n, dtype=np.int32)
    rm = np.zeros(n)
    text = C.create_string_buffer(1 << 20)
    arr = lambda lst: None if lst is None else (C.c_char_p * n)(*[s.encode() for s in lst])
    mol = None if molids is None else np.ascontiguousarray(molids, dtype=np.int32)
    rc = load().gch_lovo(dcd_path.encode(), _d(ref), n, arr(names), arr(chains), _i(mol), atom_name.encode(), cpus,
                         C.c_double(less_than_rmsd), minimum_n, write_traj.encode(), dev, rank, nranks, comm, C.byref(N),
                         _i(nat), _d(rm), C.byref(nr), C.byref(frames), C.byref(it), text, C.c_size_t(len(text)))
    if rc:
        raise _err(rc)
    lines = text.value.decode().split("\n")
    return dict(N=N.value, Natoms=nat[: N.value].copy(), RMSD=rm[: nr.value].copy(), FramesRead=frames.value,
                Iterations=it.value, PyMOLSel=lines[0], GMX=lines[1], String=lines[2] if len(lines) > 2 else "")


def LOVOResident(traj, ref, nselected_total=None, cpus=8, less_than_rmsd=1.0, minimum_n=10, comm=None, rank=0, nranks=1):
    """align.LOVO's iteration over a _lib.DeviceTraj that already holds this rank's frames."""
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    n = ref.shape[0]
    nlocal = traj.nframes
    total = nlocal if nselected_total is None else nselected_total
    N, it = C.c_int(0), C.c_int(0)
    nat = np.zeros(n, dtype=np.int32)
    rm = np.zeros(n)
    fn = load().gch_lovo_resident
    fn.argtypes = [C.c_void_p, C.c_long, C.c_long, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p,
                   C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    rc = fn(traj.h, nlocal, total, _d(ref), n, cpus, less_than_rmsd, minimum_n, comm.h if comm is not None else None, rank, nranks,
            C.byref(N), _i(nat), _d(rm), C.byref(it))
    if rc:
        raise _err(rc)
    return dict(N=N.value, Natoms=nat[: N.value].copy(), RMSD=rm.copy(), FramesRead=total, Iterations=it.value)


# ---- solv (solv/solvation.go) ----
def MDFFromCDF(cdf_sum, framesread, A, B, step):
    ret = np.array(cdf_sum, dtype=np.float64, copy=True)
    ret2 = np.zeros_like(ret)
    rc = load().gch_mdf_from_cdf(_d(ret), ret.size, int(framesread), C.c_double(A), C.c_double(B), C.c_double(step), _d(ret2))
    if rc:
        raise _err(rc)
    return ret, ret2


def EllipsoidAxes(coords, masses=None):
    c = np.ascontiguousarray(coords, dtype=np.float64)
    m = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    out = np.zeros(3)
    rc = load().gch_ellipsoid_axes(_d(c), c.shape[0], _d(m), _d(out))
    if rc:
        raise _err(rc)
    return out


def _cstrings(items):
    arr = (C.c_char_p * len(items))(*[s.encode() for s in items])
    return arr


def ConcMolRDF(dcd_path, coords0, molnames, molids, chains, masses, refindexes, residues, com=False, cpus=1, step=0.1, end=10.0,
               dev=0, rank=0, nranks=1, comm=None, window_frames=0):
    """solv.ConcMolRDF (solvation.go:140-218) on a DCD file -> (rdf, cumulative, frames read)."""
    c0 = np.ascontiguousarray(coords0, dtype=np.float64)
    n = c0.shape[0]
    ids = np.ascontiguousarray(molids, dtype=np.int32)
    ms = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    ref = np.ascontiguousarray(refindexes, dtype=np.int32)
    steps = int(end / step)
    rdf, cum = np.zeros(steps), np.zeros(steps)
    fr = C.c_int(0)
    rc = load().gch_conc_mol_rdf(dcd_path.encode(), _d(c0), n, _cstrings(list(molnames)), _i(ids), _cstrings(list(chains)), _d(ms),
                                 _i(ref), ref.size, _cstrings(list(residues)), len(residues), int(com), int(cpus), C.c_double(step),
                                 C.c_double(end), C.c_long(window_frames), dev, rank, nranks, comm, _d(rdf), _d(cum), C.byref(fr))
    if rc:
        raise _err(rc)
    return rdf, cum, fr.value


# ---- streamed variants: bounded memory, the file goes through a device window (config 5 through the reference's API) ----
def RMSDTrajStreamed(dcd_path, ref, indexes, cpus, window_frames, begin=0, skip=0, write_traj="", dev=0, rank=0, nranks=1, comm=None):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    idx = np.ascontiguousarray(indexes, dtype=np.int32)
    out = np.zeros(ref.shape[0])
    frames = C.c_int(0)
    rc = load().gch_rmsd_traj_streamed(dcd_path.encode(), _d(ref), ref.shape[0], _i(idx), idx.size, cpus, begin, skip,
                                       write_traj.encode(), C.c_long(window_frames), dev, rank, nranks, comm, _d(out), C.byref(frames))
    if rc:
        raise _err(rc)
    return out, frames.value


def LOVOStreamed(dcd_path, ref, cpus, window_frames, less_than_rmsd=1.0, minimum_n=10, dev=0):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    n = ref.shape[0]
    N, nr, fr, it = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    nat, rm = np.zeros(n, dtype=np.int32), np.zeros(n)
    rc = load().gch_lovo_streamed(dcd_path.encode(), _d(ref), n, cpus, C.c_double(less_than_rmsd), minimum_n, C.c_long(window_frames), dev,
                                  C.byref(N), _i(nat), _d(rm), C.byref(nr), C.byref(fr), C.byref(it))
    if rc:
        raise _err(rc)
    return dict(N=N.value, Natoms=nat[: N.value].copy(), RMSD=rm[: nr.value].copy(), FramesRead=fr.value, Iterations=it.value)


# ---- traj/xtc over libxdrfile (the third-party codec the reference links; bound with dlopen, GOCHEM_XDRFILE names it) ----
def xtc_available():
    return bool(load().gch_xtc_available())


def xtc_write(path, frames):
    f = np.ascontiguousarray(frames, dtype=np.float64)
    rc = load().gch_xtc_write(path.encode(), _d(f), C.c_long(f.shape[0]), f.shape[1])
    if rc:
        raise _err(rc)


def xtc_read(path, cap_frames, natoms, raw=False):
    out = np.zeros((cap_frames, natoms, 3))
    nf, na = C.c_long(0), C.c_int(0)
    rc = load().gch_xtc_read(path.encode(), _d(out), C.c_long(cap_frames), C.byref(nf), C.byref(na), int(raw))
    if rc:
        raise _err(rc)
    return out[: min(nf.value, cap_frames)], nf.value, na.value


# ---- traj/stf: goChem's own text format inside a compressed stream (zlib for .stz / .str, libzstd through dlopen for .stf) ----
def stf_zstd_available():
    return bool(load().gch_stf_zstd_available())


def stf_write(path, frames, level=6, header=None, box=None):
    f = np.ascontiguousarray(frames, dtype=np.float64)
    text = None if not header else "".join(f"{k}={v}\n" for k, v in header.items()).encode()
    b = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
    rc = load().gch_stf_write(path.encode(), _d(f), C.c_long(f.shape[0]), f.shape[1], level, text, _d(b) if b is not None else None)
    if rc:
        raise _err(rc)


def stf_read(path, cap_frames, natoms, raw=False, want_box=False):
    """-> (frames, frames in the file, atoms per frame, header dict[, box of the last frame])"""
    out = np.zeros((cap_frames, natoms, 3))
    nf, na = C.c_long(0), C.c_int(0)
    head = C.create_string_buffer(4096)
    box = np.zeros(9)
    rc = load().gch_stf_read(path.encode(), _d(out), C.c_long(cap_frames), C.byref(nf), C.byref(na), int(raw), head, C.c_size_t(4096),
                             _d(box) if want_box else None)
    if rc:
        raise _err(rc)
    header = dict(line.split("=", 1) for line in head.value.decode().splitlines() if "=" in line)
    res = (out[: min(nf.value, cap_frames)], nf.value, na.value, header)
    return res + (box,) if want_box else res


# ---- chem.Rhos / RhoShapeIndexes / EasyShape (geometric.go:511-545, handy.go:103-129) ----
def rhos_shape(moment):
    m = np.ascontiguousarray(moment, dtype=np.float64)
    rhos = np.zeros(3)
    lin, circ = C.c_double(0), C.c_double(0)
    rc = load().gch_rhos_shape(_d(m), _d(rhos), C.byref(lin), C.byref(circ))
    if rc:
        raise _err(rc)
    return rhos, lin.value, circ.value


def EasyShape(coords, masses=None):
    c = np.ascontiguousarray(coords, dtype=np.float64)
    m = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    lin, circ = C.c_double(0), C.c_double(0)
    rc = load().gch_easy_shape(_d(c), c.shape[0], _d(m) if m is not None else None, C.byref(lin), C.byref(circ))
    if rc:
        raise _err(rc)
    return lin.value, circ.value


def EasyShapeTraj(traj, masses=None, first=0, nframes=None):
    """EasyShape of every frame of a _lib.DeviceTraj -> (linear[], circular[])."""
    n = traj.nframes - first if nframes is None else nframes
    m = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    lin, circ = np.zeros(n), np.zeros(n)
    fn = load().gch_easy_shape_traj
    fn.argtypes = [C.c_void_p, C.c_long, C.c_long, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rc = fn(traj.h, first, n, _d(m) if m is not None else None, 0 if m is None else m.size, _d(lin), _d(circ))
    if rc:
        raise _err(rc)
    return lin, circ
