"""ctypes loader of the C-ABI library (include/gochem_b200.h).

Plumbing for tests/ and bench.py: numpy arrays in, numpy arrays out.  It never falls back to a CPU
implementation; a missing library or a failing CUDA call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCB_LIB") or os.path.join(HERE, "lib", "libgochem_b200.so")   # GCB_LIB: tuning builds (tools/)
HEADER = os.path.join(HERE, "..", "include", "gochem_b200.h")

F64_AOS, F32_AOS, F32_SOA, I32_AOS = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_CLUSTER = 0, 1, 2
OK, ERR_SHAPE, ERR_INDEXES, ERR_SVD, ERR_ALLOC, ERR_CUDA, ERR_ARG, EOF = 0, -1, -2, -4, -5, -7, -8, 11


class GcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gochem_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def declared_symbols() -> list[str]:
    """Every function name include/gochem_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", txt)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GcbError(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    P, I, L, D = C.c_void_p, C.c_int, C.c_long, C.c_double
    sig = {
        "gcb_version": (I, []),
        "gcb_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
        "gcb_device_count": (I, [C.POINTER(I)]),
        "gcb_device_info": (I, [I, C.POINTER(I), C.POINTER(C.c_size_t), C.POINTER(I), C.POINTER(I)]),
        "gcb_set_kernel_variant": (I, [I]),
        "gcb_set_cluster_tuning": (I, [I, I, I, I]),
        "gcb_traj_last_cluster_plan": (I, [P, C.POINTER(I)]),
        "gcb_traj_set_kernel_variant": (I, [P, I]),
        "gcb_traj_set_cluster_tuning": (I, [P, I, I, I, I]),
        "gcb_traj_set_rmsd_mode": (I, [P, I]),
        "gcb_traj_set_split_thresholds": (I, [P, I, I]),
        "gcb_set_rmsd_mode": (I, [I]),
        "gcb_set_split_thresholds": (I, [I, I]),
        "gcb_pinned_alloc": (I, [C.c_size_t, C.POINTER(P)]),
        "gcb_pinned_free": (I, [P]),
        "gcb_bind_host_thread": (I, [I]),
        "gcb_traj_create": (I, [I, I, L, C.POINTER(P)]),
        "gcb_traj_destroy": (I, [P]),
        "gcb_traj_natoms": (I, [P]),
        "gcb_traj_capacity": (L, [P]),
        "gcb_traj_nframes": (L, [P]),
        "gcb_traj_set_nframes": (I, [P, L]),
        "gcb_traj_upload": (I, [P, L, P, L, I, D]),
        "gcb_traj_upload_async": (I, [P, L, P, L, I, D, I]),
        "gcb_traj_wait": (I, [P, I]),
        "gcb_traj_upload_records_async": (I, [P, L, P, L, L, L, L, L, D, I]),
        "gcb_traj_stage_bytes": (C.c_size_t, [P]),
        "gcb_traj_download": (I, [P, L, L, P]),
        "gcb_super_rmsd": (I, [P, L, L, P, I, P, P, I, I, P, P, P, P]),
        "gcb_super_rmsd_async": (I, [P, L, L, P, I, P, P, I, I]),
        "gcb_fetch_results": (I, [P, L, L, P, P, P, P]),
        "gcb_rmsd": (I, [P, L, L, P, I, P, P, I, P]),
        "gcb_peratom_dist": (I, [P, L, P, I, P, P, I, P]),
        "gcb_peratom_dev_sum": (I, [P, L, L, P, P, I, I, P, P]),
        "gcb_lovo_resident": (I, [P, L, L, P, P, I, D, D, I, I, P, C.POINTER(I), C.POINTER(I), C.POINTER(I), P, P, P]),
        "gcb_allpairs_rmsd": (I, [P, L, L, P, I, L, L, P, C.POINTER(C.c_float), C.POINTER(L)]),
        "gcb_allpairs_rmsd_fit": (I, [P, L, L, P, I, L, L, P, C.POINTER(C.c_float), C.POINTER(L)]),
        "gcb_moments": (I, [P, L, L, P, P, I, P, P]),
        "gcb_rdf_cdf_sum": (I, [P, L, L, P, P, P, P, I, I, D, D, I, P, P, C.POINTER(L)]),
        "gcb_moments_async": (I, [P, L, L, P, P, I]),
        "gcb_moments_fetch": (I, [P, L, P, P]),
        "gcb_timer_start": (I, [P]),
        "gcb_timer_stop": (I, [P, C.POINTER(C.c_float)]),
        "gcb_sync": (I, [P]),
        "gcb_launch_count": (L, []),
        "gcb_synth_fill": (I, [P, L, L, P, P, C.c_ulonglong, L, I, I]),
        "gcb_comm_unique_id": (I, [C.c_char_p]),
        "gcb_comm_init": (I, [I, I, I, C.c_char_p, C.POINTER(P)]),
        "gcb_comm_destroy": (I, [P]),
        "gcb_allreduce_sum_f64": (I, [P, P, L]),
        "gcb_traj_allgather": (I, [P, P, C.POINTER(L)]),
        "gcb_allpairs_rmsd_sharded": (I, [P, L, L, P, I, I, P, P, C.POINTER(L), C.POINTER(L), C.POINTER(C.c_float),
                                          C.POINTER(C.c_float), C.POINTER(L)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    load().gcb_last_error(buf, 2048)
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != 0:
        raise GcbError(rc, last_error())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _idx(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().gcb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class PinnedBuffer:
    """C-owned page-locked staging buffer exposed as a numpy array."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(load().gcb_pinned_alloc(nbytes, C.byref(self.ptr)))
        self.nbytes = nbytes
        self.u8 = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def view(self, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self.u8[:n].view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            load().gcb_pinned_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Comm:
    def __init__(self, dev: int, nranks: int, rank: int, unique_id: bytes):
        self.h = C.c_void_p()
        check(load().gcb_comm_init(dev, nranks, rank, unique_id, C.byref(self.h)))
        self.nranks, self.rank = nranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(load().gcb_comm_unique_id(buf))
        return buf.raw

    def allreduce_sum(self, vec: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(vec, dtype=np.float64).copy()
        check(load().gcb_allreduce_sum_f64(self.h, _ptr(v), v.size))
        return v

    def close(self):
        if self.h:
            load().gcb_comm_destroy(self.h)
            self.h = C.c_void_p()


class DeviceTraj:
    """Device-resident trajectory (the batch of v3 device matrices)."""

    def __init__(self, natoms: int, cap_frames: int, dev: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        check(self.lib.gcb_traj_create(dev, natoms, cap_frames, C.byref(self.h)))
        self.natoms, self.cap, self.dev = natoms, cap_frames, dev

    # --- settings of this handle (the gcb_set_* forms only change the defaults of handles created later) ---
    def set_kernel_variant(self, variant: int):
        check(self.lib.gcb_traj_set_kernel_variant(self.h, variant))

    def set_cluster_tuning(self, cluster=0, stages=0, depth=0, resident=-1):
        check(self.lib.gcb_traj_set_cluster_tuning(self.h, cluster, stages, depth, resident))

    def set_rmsd_mode(self, explicit_only: int):
        check(self.lib.gcb_traj_set_rmsd_mode(self.h, int(explicit_only)))

    def set_split_thresholds(self, write_back_max_atoms=-1, per_atom_max_atoms=-1):
        check(self.lib.gcb_traj_set_split_thresholds(self.h, write_back_max_atoms, per_atom_max_atoms))

    def last_cluster_plan(self) -> dict:
        out = (C.c_int * 8)()
        check(self.lib.gcb_traj_last_cluster_plan(self.h, out))
        keys = ["cluster", "part", "atoms_per_thread", "stages", "depth", "resident", "clusters", "smem"]
        d = dict(zip(keys, list(out)))
        d["consumer_warps"], d["resident"] = d["resident"] >> 4, d["resident"] & 1
        return d

    # --- data movement ---
    def upload(self, frames: np.ndarray, first: int = 0, layout: int = F64_AOS, scale: float = 1.0):
        dt = np.float64 if layout == F64_AOS else np.int32 if layout == I32_AOS else np.float32
        a = np.ascontiguousarray(frames, dtype=dt)
        n = a.size // (3 * self.natoms)
        check(self.lib.gcb_traj_upload(self.h, first, _ptr(a), n, layout, scale))
        return n

    def upload_async(self, pinned_view: np.ndarray, nframes: int, first: int, layout: int, scale: float, slot: int):
        check(self.lib.gcb_traj_upload_async(self.h, first, _ptr(pinned_view), nframes, layout, scale, slot))

    def wait(self, slot: int):
        check(self.lib.gcb_traj_wait(self.h, slot))

    def download(self, first: int = 0, nframes: int | None = None) -> np.ndarray:
        n = self.nframes - first if nframes is None else nframes
        out = np.empty((n, self.natoms, 3), dtype=np.float64)
        check(self.lib.gcb_traj_download(self.h, first, n, _ptr(out)))
        return out

    @property
    def nframes(self) -> int:
        return self.lib.gcb_traj_nframes(self.h)

    def set_nframes(self, n: int):
        check(self.lib.gcb_traj_set_nframes(self.h, n))

    def synth_fill(self, ref, sigma, nframes, first=0, seed=12347, frame_id0=0, frame0_is_ref=False, through_f32=False):
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        check(self.lib.gcb_synth_fill(self.h, first, nframes, _ptr(ref), _ptr(sigma), seed, frame_id0,
                                      int(frame0_is_ref), int(through_f32)))

    # --- compute ---
    def super_rmsd(self, ref, idx_test=None, idx_ref=None, write_back=False, first=0, nframes=None,
                   want=("rmsd", "rot", "trans1", "trans2")):
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        it, ir = _idx(idx_test), _idx(idx_ref)
        nidx = 0 if it is None else it.size
        out = {
            "rmsd": np.empty(n) if "rmsd" in want else None,
            "rot": np.empty((n, 3, 3)) if "rot" in want else None,
            "trans1": np.empty((n, 3)) if "trans1" in want else None,
            "trans2": np.empty((n, 3)) if "trans2" in want else None,
        }
        check(self.lib.gcb_super_rmsd(self.h, first, n, _ptr(ref), ref.shape[0], _ptr(it), _ptr(ir), nidx,
                                      int(write_back), _ptr(out["rmsd"]), _ptr(out["rot"]), _ptr(out["trans1"]),
                                      _ptr(out["trans2"])))
        return out

    def super_rmsd_async(self, ref, idx_test=None, idx_ref=None, write_back=False, first=0, nframes=None):
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        it, ir = _idx(idx_test), _idx(idx_ref)
        nidx = 0 if it is None else it.size
        check(self.lib.gcb_super_rmsd_async(self.h, first, n, _ptr(ref), ref.shape[0], _ptr(it), _ptr(ir), nidx,
                                            int(write_back)))

    def fetch(self, first=0, nframes=None, want=("rmsd",)):
        n = self.nframes - first if nframes is None else nframes
        out = {
            "rmsd": np.empty(n) if "rmsd" in want else None,
            "rot": np.empty((n, 3, 3)) if "rot" in want else None,
            "trans1": np.empty((n, 3)) if "trans1" in want else None,
            "trans2": np.empty((n, 3)) if "trans2" in want else None,
        }
        check(self.lib.gcb_fetch_results(self.h, first, n, _ptr(out["rmsd"]), _ptr(out["rot"]), _ptr(out["trans1"]),
                                         _ptr(out["trans2"])))
        return out

    def rmsd(self, ref, idx_test=None, idx_ref=None, first=0, nframes=None) -> np.ndarray:
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        it, ir = _idx(idx_test), _idx(idx_ref)
        nidx = 0 if it is None else it.size
        out = np.empty(n)
        check(self.lib.gcb_rmsd(self.h, first, n, _ptr(ref), ref.shape[0], _ptr(it), _ptr(ir), nidx, _ptr(out)))
        return out

    def peratom_dev_sum(self, ref, idx=None, write_back=False, first=0, nframes=None, comm: Comm | None = None):
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        it = _idx(idx)
        nidx = 0 if it is None else it.size
        out = np.empty(self.natoms)
        check(self.lib.gcb_peratom_dev_sum(self.h, first, n, _ptr(ref), _ptr(it), nidx, int(write_back),
                                           comm.h if comm else None, _ptr(out)))
        return out

    def peratom_dev_sum_atoms(self, ref, idx, want, first=0, nframes=None, comm: Comm | None = None):
        """Per-atom sums for the atoms of `want` only (align.LOVO's candidates)."""
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        it, iw = _idx(idx), _idx(want)
        out = np.empty(iw.size)
        check(self.lib.gcb_peratom_dev_sum_atoms(self.h, first, n, _ptr(ref), _ptr(it), 0 if it is None else it.size, _ptr(iw), iw.size,
                                                 comm.h if comm else None, _ptr(out)))
        return out

    def lovo_resident(self, ref, cand, total_frames=None, less_than_rmsd=1.0, minimum_n=10, max_iter=100000, first=0, nframes=None,
                      comm: Comm | None = None):
        """align.LOVO's iteration with the bookkeeping on the device (gcb_lovo_resident).  None when the library does not take
        the case (return code 1): run the host loop then."""
        n = self.nframes - first if nframes is None else nframes
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        ic = _idx(cand)
        nout, iters, dis = C.c_int(0), C.c_int(0), C.c_int(0)
        old, ids, vals = np.zeros(ic.size, dtype=np.int32), np.zeros(ic.size, dtype=np.int32), np.zeros(ic.size)
        rc = self.lib.gcb_lovo_resident(self.h, first, n, _ptr(ref), _ptr(ic), ic.size, float(n if total_frames is None else total_frames),
                                        float(less_than_rmsd), int(minimum_n), int(max_iter), comm.h if comm else None,
                                        C.byref(nout), C.byref(iters), C.byref(dis), _ptr(old), _ptr(ids), _ptr(vals))
        if rc == 1:
            return None
        check(rc)
        return dict(N=nout.value, Iterations=iters.value, Disagreement=dis.value, indexesold=old, ids=ids, vals=vals,
                    Natoms=old[: nout.value].copy())

    def allpairs_rmsd(self, idx=None, first=0, nframes=None, row0=0, nrows=None, fit=False):
        n = self.nframes - first if nframes is None else nframes
        nr = n - row0 if nrows is None else nrows
        it = _idx(idx)
        out = np.empty((nr, n))
        ms, redone = C.c_float(0), C.c_long(0)
        fn = self.lib.gcb_allpairs_rmsd_fit if fit else self.lib.gcb_allpairs_rmsd
        check(fn(self.h, first, n, _ptr(it), 0 if it is None else it.size, row0, nr, _ptr(out), C.byref(ms), C.byref(redone)))
        return out, ms.value, redone.value

    def allgather(self, comm: "Comm", counts):
        """Frames sharded over the ranks (each rank's block already at its offset) -> all frames on every rank."""
        c = (C.c_long * len(counts))(*[int(x) for x in counts])
        check(self.lib.gcb_traj_allgather(self.h, comm.h, c))

    def allpairs_rmsd_sharded(self, comm: "Comm", idx=None, first=0, nframes=None, fit=False, fetch=True):
        """Multi-GPU all-pairs RMSD matrix: returns (this rank's row block or None, row0, gemm_ms, total_ms, redone)."""
        n = self.nframes - first if nframes is None else nframes
        it = _idx(idx)
        base, rem = divmod(n, comm.nranks)
        nr = base + (1 if comm.rank < rem else 0)
        out = np.empty((nr, n)) if fetch else None
        row0, nrows, redone = C.c_long(0), C.c_long(0), C.c_long(0)
        gms, tms = C.c_float(0), C.c_float(0)
        check(self.lib.gcb_allpairs_rmsd_sharded(self.h, first, n, _ptr(it), 0 if it is None else it.size, 1 if fit else 0, comm.h,
                                                 _ptr(out), C.byref(row0), C.byref(nrows), C.byref(gms), C.byref(tms), C.byref(redone)))
        assert nrows.value == nr
        return out, row0.value, gms.value, tms.value, redone.value

    def moments(self, mass=None, idx=None, first=0, nframes=None):
        """chem.CenterOfMass + chem.MomentTensor for every frame -> (F x 3, F x 3 x 3)."""
        n = self.nframes - first if nframes is None else nframes
        m = None if mass is None else np.ascontiguousarray(mass, dtype=np.float64)
        it = _idx(idx)
        cen, mom = np.empty((n, 3)), np.empty((n, 3, 3))
        check(self.lib.gcb_moments(self.h, first, n, _ptr(m), _ptr(it), 0 if it is None else it.size, _ptr(cen), _ptr(mom)))
        return cen, mom

    def rdf_cdf_sum(self, molid, chain, flags, refidx, com=False, step=0.1, end=10.0, cpus=1, comm=None, first=0, nframes=None):
        """Frame loop of solv.ConcMolRDF -> (summed cumulative counts per shell, frames read)."""
        n = self.nframes - first if nframes is None else nframes
        m = np.ascontiguousarray(molid, dtype=np.int32)
        c = np.ascontiguousarray(chain, dtype=np.int32)
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        r = _idx(refidx)
        out = np.zeros(int(end / step))
        read = C.c_long(0)
        check(self.lib.gcb_rdf_cdf_sum(self.h, first, n, _ptr(m), _ptr(c), _ptr(f), _ptr(r), r.size, int(com), step, end, cpus,
                                       comm.h if comm is not None else None, _ptr(out), C.byref(read)))
        return out, read.value

    def moments_async(self, mass=None, idx=None, first=0, nframes=None):
        n = self.nframes - first if nframes is None else nframes
        m = None if mass is None else np.ascontiguousarray(mass, dtype=np.float64)
        it = _idx(idx)
        check(self.lib.gcb_moments_async(self.h, first, n, _ptr(m), _ptr(it), 0 if it is None else it.size))

    # --- timing on the library's stream ---
    def timer_start(self):
        check(self.lib.gcb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.lib.gcb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def sync(self):
        check(self.lib.gcb_sync(self.h))

    def close(self):
        if self.h:
            self.lib.gcb_traj_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
