"""CPU tests of the host-side mirror of the reference's Go API: LOVO bookkeeping, frame selection, sharding,
the DCD reader / writer.  The trajectory pass itself is GPU-only; here the per-atom means fed to the LOVO state
machine come from the oracle, so the bookkeeping is checked against oracle.lovo end to end."""
import os
import struct

import numpy as np
import pytest

import oracle
from gochem_b200 import host_api as H
from gochem_b200 import synth
from oracle import oracle_np as onp


def test_host_library_exports_every_declared_symbol():
    lib = H.load()
    names = H.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_mobile_limit_matches_oracle():
    rng = np.random.default_rng(2)
    for _ in range(200):
        vals = np.sort(rng.uniform(0.1, 4.0, int(rng.integers(12, 60))))
        lim = float(rng.uniform(0.05, 3.0))
        tol = int(rng.integers(0, 12))
        assert H.mobile_limit(lim, tol, vals) == oracle.mobile_limit(lim, tol, vals)
    with pytest.raises(H.GoChemError):
        H.mobile_limit(1.0, 10, np.array([0.1, 2.0]))  # the reference would loop forever


def test_selected_frames_follow_the_reference_quirks():
    # chunks of cpus, tail dropped (lovo.go:323-326)
    assert list(H.selected_frames(11, 4)) == list(range(8))
    assert list(H.selected_frames(3, 4)) == []
    # begin: begin/cpus chunks are "skipped" (:306), but each skipped chunk is followed, in the same loop turn, by a
    # chunk that IS processed -- there is no `continue` (:316-322)
    assert list(H.selected_frames(20, 4, begin=8)) == [4, 5, 6, 7, 12, 13, 14, 15, 16, 17, 18, 19]
    assert list(H.selected_frames(20, 4, begin=7)) == list(range(4, 20))
    # skip: a discarded chunk is followed by a processed one, no `continue` (:316-322)
    assert list(H.selected_frames(24, 2, skip=2)) == [0, 1, 4, 5, 6, 7, 10, 11, 12, 13, 16, 17, 18, 19, 22, 23]
    # against the oracle's chunk loop for random settings
    rng = np.random.default_rng(3)
    ref, frames = synth.make_case(5, 40)
    for _ in range(30):
        F, cpus = int(rng.integers(1, 40)), int(rng.integers(1, 7))
        begin, skip = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        _, total = oracle.rmsd_traj(frames[:F], ref, np.arange(5), cpus, begin, skip)
        assert len(H.selected_frames(F, cpus, begin, skip)) == total


def test_shard_ranges_partition_the_selection():
    for nsel in [0, 1, 7, 64, 1001]:
        for nranks in [1, 2, 3, 8]:
            got = []
            for r in range(nranks):
                f, c = H.shard_range(nsel, r, nranks)
                got += list(range(f, f + c))
                assert c in (nsel // nranks, nsel // nranks + 1)
            assert got == list(range(nsel))


def drive_lovo(frames, ref, full, cpus, less=1.0, minn=10, allreduce=None, shard=None):
    """align.LOVO with the host bookkeeping from the product and the trajectory pass from the oracle."""
    st = H.LovoState(full, less, minn)
    sel = H.selected_frames(frames.shape[0], cpus)
    used = frames[sel]
    while True:
        idx = st.fit_indexes()
        if shard is None:
            sums, total = oracle.rmsd_traj(used, ref, idx, cpus=1)
            sums = sums * total
        else:
            f, c = shard
            sums, total = oracle.rmsd_traj(used[f:f + c], ref, idx, cpus=1) if c else (np.zeros(ref.shape[0]), 0)
            sums = allreduce(np.nan_to_num(sums) * total)
        if st.step(sums / len(sel)):
            break
    return st.finish(ref.shape[0], len(sel))


def test_lovo_state_machine_matches_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "lovo_n150_f42.npz"))
    ref, frames, cpus = z["ref"].astype(np.float64), z["frames"].astype(np.float64), int(z["cpus"])
    got = drive_lovo(frames, ref, np.arange(150), cpus)
    assert got["N"] == int(z["N"]) and got["Iterations"] == int(z["Iterations"]) and got["FramesRead"] == 40
    assert list(got["Natoms"]) == list(z["Natoms"])
    assert np.abs(got["RMSD"] - z["RMSD"]).max() < 1e-11
    # a second, larger case against the C oracle, candidates = every third atom
    ref, frames = synth.make_case(240, 60)
    full = np.arange(0, 240, 3)
    want = oracle.lovo(frames, ref, full, cpus=5, less_than_rmsd=0.8, minimum_n=7)
    got = drive_lovo(frames, ref, full, 5, 0.8, 7)
    assert got["N"] == want["N"] and got["Iterations"] == want["Iterations"]
    assert list(got["Natoms"]) == list(want["Natoms"])
    assert np.abs(got["RMSD"] - want["RMSD"]).max() < 1e-11
    # LessThanRMSD <= 0: n stays at the full candidate list
    want = oracle.lovo(frames, ref, full, cpus=5, less_than_rmsd=0.0)
    got = drive_lovo(frames, ref, full, 5, 0.0, 10)
    assert got["N"] == want["N"] == 80 and got["Iterations"] == want["Iterations"]


def test_dcd_round_trip(tmp_path):
    ref, frames = synth.make_case(37, 9, through_f32=True)
    p = str(tmp_path / "t.dcd")
    H.dcd_write(p, frames)
    raw = open(p, "rb").read()
    # header bytes the reference's writer emits (dcd_write.go:102-205) and the running frame counter (:300-330)
    assert struct.unpack("<i4si", raw[:12]) == (84, b"CORD", 9)
    assert struct.unpack("<i", raw[84:88])[0] == 24 and struct.unpack("<i", raw[88:92])[0] == 84
    assert struct.unpack("<f", raw[44:48])[0] == 1.0
    natoms_off = 92 + 4 + 4 + 160 + 4 + 4
    assert struct.unpack("<i", raw[natoms_off:natoms_off + 4])[0] == 37
    assert len(raw) == natoms_off + 8 + 9 * 3 * (37 * 4 + 8)
    back, nf, na = H.dcd_read(p, 20, 37)
    assert (nf, na) == (9, 37) and np.array_equal(back, frames)
    # NextConc in chunks of 4: the tail of 1 is dropped
    back, nf, na = H.dcd_read(p, 20, 37, chunk=4)
    assert nf == 8 and np.array_equal(back, frames[:8])
    with pytest.raises(H.GoChemError):
        H.dcd_read(str(tmp_path / "missing.dcd"), 1, 1)
    # big-endian files are read too (dcd.go:172-174)
    be = bytearray(raw)
    for off in list(range(0, 4, 4)) + list(range(8, 92, 4)) + [92, 96, 260, 264, 268, 272]:
        be[off:off + 4] = be[off:off + 4][::-1]
    body = natoms_off + 8
    for off in range(body, len(raw), 4):
        be[off:off + 4] = be[off:off + 4][::-1]
    pb = str(tmp_path / "be.dcd")
    open(pb, "wb").write(bytes(be))
    back, nf, na = H.dcd_read(pb, 20, 37)
    assert nf == 9 and np.array_equal(back, frames)


def test_xtc_reader_over_libxdrfile(tmp_path):
    """traj/xtc mirror: XTCObj.Next (x10, widened per element, xtc.go:129-134) and the raw float32 feed the device staging uses,
    on a file written with the codec's own writer.  The codec is the third-party libxdrfile the reference links where it is
    installed, tests/xdrshim elsewhere (conftest.py)."""
    assert H.xtc_available()
    rng = np.random.default_rng(0)
    frames = rng.normal(0, 15, (9, 61, 3))
    p = str(tmp_path / "t.xtc")
    H.xtc_write(p, frames)
    a, nf, na = H.xtc_read(p, 20, 61)
    assert (nf, na) == (9, 61)
    assert np.abs(a - frames).max() < 0.006          # precision 1000 per nm: 0.01 A steps
    b, nf2, _ = H.xtc_read(p, 20, 61, raw=True)
    assert nf2 == 9 and np.array_equal(a, b)          # 10 * float64(float32) either way
    with pytest.raises(H.GoChemError):
        H.xtc_read(str(tmp_path / "missing.xtc"), 1, 1)


def test_xtc_reader_over_the_shim_in_a_fresh_process(tmp_path):
    """The same reader bound to tests/xdrshim (what the GPU box uses, where libxdrfile is absent).  The binding is made once per
    process, hence the subprocess.  The shim stores float32 nanometres uncompressed, so the round trip is exact to float32."""
    import subprocess
    import sys

    from conftest import ROOT, build_xdrshim

    code = (
        "import numpy as np\n"
        "from gochem_b200 import host_api as H\n"
        "assert H.xtc_available()\n"
        "rng = np.random.default_rng(0)\n"
        "frames = rng.normal(0, 15, (9, 61, 3))\n"
        "H.xtc_write(r'%s', frames)\n"
        "a, nf, na = H.xtc_read(r'%s', 20, 61)\n"
        "assert (nf, na) == (9, 61)\n"
        "want = 10.0 * (frames / 10.0).astype(np.float32).astype(np.float64)\n"
        "assert np.array_equal(a, want), np.abs(a - want).max()\n"
        "b, nf2, _ = H.xtc_read(r'%s', 20, 61, raw=True)\n"
        "assert nf2 == 9 and np.array_equal(a, b)\n"
        "print('shim ok')\n"
    ) % ((str(tmp_path / "s.xtc"),) * 3)
    env = dict(os.environ, GOCHEM_XDRFILE=build_xdrshim(), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "shim ok" in r.stdout, r.stderr[-2000:]


def test_rotator_shape_error_comes_from_the_library():
    """geometric.go:130-132: different row counts -> "goChem: Ill-formed matrices", raised by the host library itself (the check
    precedes any device work, so it runs without a GPU)."""
    a, b = np.zeros((5, 3)), np.zeros((4, 3))
    with pytest.raises(H.GoChemError, match="Ill-formed matrices") as e:
        H.RotatorTranslatorToSuper(a, b)
    assert "RotatorTranslatorToSuper" in str(e.value)


def test_dcd_bulk_record_feed(tmp_path):
    """ConcTraj::NextRecords: whole runs of raw frames in one read, every window size, against the frame-by-frame reader; also on
    CHARMM files that carry a unit-cell record before every frame and a fourth-dimension record after it (dcd.go:361-376, 404-417)."""
    natoms, nframes = 37, 11
    ref, frames = synth.make_case(natoms, nframes, through_f32=True)
    p = str(tmp_path / "plain.dcd")
    H.dcd_write(p, frames)
    for win in (1, 4, 11, 1000):
        back, nf, na = H.dcd_read(p, 20, natoms, chunk=-win)
        assert (nf, na) == (nframes, natoms) and np.array_equal(back, frames)
    raw = open(p, "rb").read()
    body = 92 + 4 + 4 + 160 + 4 + 4 + 8
    plane = 4 * natoms
    rec = 3 * (plane + 8)
    assert len(raw) == body + nframes * rec
    cell = struct.pack("<i6di", 48, 50.0, 90.0, 50.0, 90.0, 90.0, 50.0, 48)
    for flag in (2, 1):   # 2: unit cell only; 1: the reference then also expects a fourth-dimension record (dcd.go:196-201)
        out = bytearray(raw[:body])
        out[48:52] = struct.pack("<i", flag)
        for f in range(nframes):
            out += cell + raw[body + f * rec: body + (f + 1) * rec]
            if flag == 1:
                out += struct.pack("<i", plane) + bytes(plane) + struct.pack("<i", plane)
        pc = str(tmp_path / f"charmm{flag}.dcd")
        open(pc, "wb").write(bytes(out))
        slow, nf, na = H.dcd_read(pc, 20, natoms)
        assert nf == nframes and np.array_equal(slow, frames)
        for win in (1, 3, 1000):
            back, nf, na = H.dcd_read(pc, 20, natoms, chunk=-win)
            assert nf == nframes and np.array_equal(back, frames)
    # a file cut in the middle of a frame: the partial frame is not delivered
    pt = str(tmp_path / "cut.dcd")
    open(pt, "wb").write(raw[: body + 5 * rec + 40])
    back, nf, na = H.dcd_read(pt, 20, natoms, chunk=-4)
    assert nf == 5 and np.array_equal(back, frames[:5])


def _gloo_worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref, frames = synth.make_case(90, 33)
    full = np.arange(90)
    cpus = 4
    nsel = len(H.selected_frames(frames.shape[0], cpus))

    def allreduce(v):
        # all-gather + rank-ordered sum: what gcb_allreduce_sum_f64 does with NCCL
        t = torch.from_numpy(np.ascontiguousarray(v))
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        out = torch.zeros_like(t)
        for p in parts:
            out += p
        return out.numpy()

    got = drive_lovo(frames, ref, full, cpus, allreduce=allreduce, shard=H.shard_range(nsel, rank, world))
    np.savez(os.path.join(tmpdir, f"r{rank}.npz"), N=got["N"], Natoms=got["Natoms"], RMSD=got["RMSD"], It=got["Iterations"])
    dist.destroy_process_group()


def test_sharded_lovo_two_ranks_gloo(tmp_path):
    """N > 1 path on CPU: frames sharded over 2 ranks, per-atom sums combined with a gloo all-gather in rank
    order; both ranks must reach the single-rank answer with identical index lists."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref, frames = synth.make_case(90, 33)
    want = oracle.lovo(frames, ref, np.arange(90), cpus=4)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for r in (r0, r1):
        assert int(r["N"]) == want["N"] and int(r["It"]) == want["Iterations"]
        assert list(r["Natoms"]) == list(want["Natoms"])
        assert np.abs(r["RMSD"] - want["RMSD"]).max() < 1e-12
    assert np.array_equal(r0["RMSD"], r1["RMSD"])  # bit-identical across ranks


def test_kernel_rotation_solver_against_lapack():
    """csrc/kabsch3.cuh (the 3x3 solve the CUDA kernels run, compiled for the host here) against LAPACK gesvd with the
    reference's reflection fix (geometric.go:174-190)."""
    import ctypes as C

    import scipy.linalg as sla

    lib = H.load()
    rng = np.random.default_rng(4)

    def solve(h, which):
        h = np.ascontiguousarray(h, dtype=np.float64)
        r = np.zeros((3, 3))
        rc = lib.gch_kabsch3(h.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double)), which)
        return rc, r

    npolar = nrefl = 0
    for k in range(3000):
        if k % 2:
            a = rng.standard_normal((40, 3)) * 12 * rng.uniform(0.1, 1, 3)
            b = a + rng.standard_normal((40, 3)) * rng.uniform(0, 3)
            if k % 10 == 1:
                b = b * np.array([1.0, 1.0, -1.0])          # mirror image: forces the reflection branch
            h = (a - a.mean(0)).T @ (b - b.mean(0))
        else:
            h = rng.standard_normal((3, 3)) * rng.uniform(1e-3, 1e6)
        u, s, vt = sla.svd(h, lapack_driver="gesvd")
        want = u @ vt
        if np.linalg.det(want) < 0:
            want = u @ np.diag([1.0, 1.0, -1.0]) @ vt
            nrefl += 1
        gap = (s[1] + s[2]) / s[0] if np.linalg.det(h) > 0 else (s[1] - s[2]) / s[0]
        if gap < 1e-3:
            continue                                          # rotation not unique enough to compare
        rc, r = solve(h, 0)
        assert rc == 0 and np.abs(r - want).max() < 1e-10 / min(1.0, gap)
        assert np.abs(r @ r.T - np.eye(3)).max() < 1e-13 and np.linalg.det(r) > 0
        rc, rj = solve(h, 1)
        assert rc == 0 and np.abs(rj - want).max() < 1e-10 / min(1.0, gap)
        rc, rp = solve(h, 2)
        if rc == 0:
            npolar += 1
            assert np.linalg.det(h) > 0 and np.abs(rp - want).max() < 1e-10 / min(1.0, gap)
    assert npolar > 1000 and nrefl > 300
    # non-finite input is the reference's "mat.SVD failed"
    assert solve(np.full((3, 3), np.nan), 0)[0] == -4


def test_mdf_from_cdf_matches_oracle():
    """solv.MDFFromCDF (solvation.go:317-352) in the host mirror against the oracle; A, B < 1 is an error in both."""
    rng = np.random.default_rng(4)
    cdf = np.cumsum(rng.integers(0, 30, 100)).astype(np.float64) * 17
    a, a2 = H.MDFFromCDF(cdf, 17, 1.7, 1.2, 0.1)
    b, b2 = oracle.mdf_from_cdf(cdf, 17, 1.7, 1.2, 0.1)
    assert np.array_equal(a, b) and np.array_equal(a2, b2)
    assert a[-1] == 1.0
    with pytest.raises(H.GoChemError):
        H.MDFFromCDF(cdf, 17, 0.9, 1.2, 0.1)


def _stf_text(frames, natoms_line=True, header=None, box=None):
    """An STF stream as the reference's writer lays it out (stf.go:64-110, 165-189): header lines, '** natoms', per frame
    one '%d %d %d' line per atom (coordinate * 100 rounded half to even) and a line starting with '*'."""
    out = []
    for k, v in (header or {}).items():
        out.append(f"{k}={v}\n")
    out.append(f"** {frames.shape[1]}\n")
    for fr in frames:
        for x, y, z in np.rint(fr * 100.0).astype(np.int64):
            out.append(f"{x} {y} {z}\n")
        out.append("*\n" if box is None else "* " + " ".join("%4.2f" % b for b in box) + "\n")
    return "".join(out).encode()


def test_stf_reader_against_files_written_by_python(tmp_path):
    """traj/stf mirror: files built here with Python's gzip / zlib (the 'z' and 'r' forms of stf.go:245-285) read back through
    StfR.Next and through the raw int32 feed of the device staging: float64(f) / 100 exactly (stf.go:341)."""
    import gzip
    import zlib

    rng = np.random.default_rng(7)
    frames = rng.normal(0, 30, (9, 37, 3))
    frames[0, 0] = [0.005, 0.015, -0.025]          # ties: rounded half to even -> 0, 2, -2 hundredths
    want = np.rint(frames * 100.0) / 100.0
    text = _stf_text(frames, header={"origin": "python", "prec": "3"}, box=[10, 0, 0, 0, 20.5, 0, 0, 0, 30])
    pz = str(tmp_path / "t.stz")
    with gzip.open(pz, "wb") as f:
        f.write(text)
    pr = str(tmp_path / "t.str")
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    with open(pr, "wb") as f:
        f.write(co.compress(text) + co.flush())
    for p in (pz, pr):
        got, nf, na, header, box = H.stf_read(p, 20, 37, want_box=True)
        assert (nf, na) == (9, 37) and np.array_equal(got, want)
        assert header == {"origin": "python", "prec": "3"}      # the "prec" entry is read and ignored (stf.go:311-318): p stays 100
        assert np.array_equal(box, [10, 0, 0, 0, 20.5, 0, 0, 0, 30])
        raw, nf2, _, _ = H.stf_read(p, 20, 37, raw=True)
        assert nf2 == 9 and np.array_equal(raw, want)
    assert want[0, 0].tolist() == [0.0, 0.02, -0.02]
    # two gzip members back to back read as one stream (gzip.Reader's default)
    pm = str(tmp_path / "multi.stz")
    cut = text.index(b"*\n") + 2 if b"*\n" in text else len(text) // 2
    with open(pm, "wb") as f:
        f.write(gzip.compress(text[: len(text) // 2]) + gzip.compress(text[len(text) // 2:]))
    got, nf, _, _ = H.stf_read(pm, 20, 37)
    assert nf == 9 and np.array_equal(got, want)


def test_stf_writer_round_trip_and_text_layout(tmp_path):
    """StfW writes what the reference writes ('%d %d %d' lines, '*' terminators, '** natoms' after the header), readable by
    Python's gzip, and the zstd form ('.stf', the default codec) round-trips through libzstd where that library loads."""
    import gzip

    rng = np.random.default_rng(8)
    frames = rng.normal(0, 25, (6, 50, 3))
    want = np.rint(frames * 100.0) / 100.0
    p = str(tmp_path / "w.stz")
    H.stf_write(p, frames, level=5, header={"a": "1", "b": "two"})
    with gzip.open(p, "rb") as f:
        assert f.read() == _stf_text(frames, header={"a": "1", "b": "two"})
    got, nf, na, header = H.stf_read(p, 10, 50)
    assert (nf, na, header) == (6, 50, {"a": "1", "b": "two"}) and np.array_equal(got, want)
    H.stf_write(str(tmp_path / "w.str"), frames, level=9)
    got, nf, _, _ = H.stf_read(str(tmp_path / "w.str"), 10, 50)
    assert nf == 6 and np.array_equal(got, want)
    # the reference's default level 11 is refused by flate / gzip there as well (stf.go:115, 131-135)
    with pytest.raises(H.GoChemError, match="invalid compression level"):
        H.stf_write(str(tmp_path / "bad.stz"), frames, level=11)
    if H.stf_zstd_available():
        pf = str(tmp_path / "w.stf")
        H.stf_write(pf, frames, level=11, box=[1, 2, 3, 4, 5, 6, 7, 8, 9])
        got, nf, _, _, box = H.stf_read(pf, 10, 50, want_box=True)
        assert nf == 6 and np.array_equal(got, want) and np.array_equal(box, np.arange(1.0, 10.0))


def test_stf_reader_errors(tmp_path):
    """Malformed streams fail where the reference fails (stf.go:324-345, 352-389): field counts, non-integers, a missing
    terminator, a stream that ends inside a frame; a missing file; the end of the trajectory is a LastFrameError (nf counts)."""
    import gzip

    def write(name, text):
        p = str(tmp_path / name)
        with gzip.open(p, "wb") as f:
            f.write(text)
        return p

    good = b"** 2\n1 2 3\n4 5 6\n*\n"
    got, nf, na, _ = H.stf_read(write("ok.stz", good), 4, 2)
    assert (nf, na) == (1, 2) and np.array_equal(got[0], [[0.01, 0.02, 0.03], [0.04, 0.05, 0.06]])
    for name, text, msg in [
        ("few.stz", b"** 2\n1 2\n4 5 6\n*\n", "Too few fields"),
        ("many.stz", b"** 2\n1 2 3 4\n4 5 6\n*\n", "Too many fields"),
        ("nan.stz", b"** 2\n1 2 x\n4 5 6\n*\n", "Can't parse coordinate 2"),
        ("float.stz", b"** 2\n1 2 3.5\n4 5 6\n*\n", "Can't parse coordinate 2"),
        ("term.stz", b"** 2\n1 2 3\n4 5 6\n7 8 9\n*\n", "Wrong number of atoms in frame"),
        ("cut.stz", b"** 2\n1 2 3\n4 5 6\n*\n1 2 3\n", "EOF"),
        ("head.stz", b"no header here\n", "Malformed header"),
        ("nat.stz", b"** many\n", "Can't read atom number"),
    ]:
        with pytest.raises(H.GoChemError, match=msg):
            H.stf_read(write(name, text), 4, 2)
    with pytest.raises(H.GoChemError, match="no such file"):
        H.stf_read(str(tmp_path / "missing.stz"), 1, 1)
    p = str(tmp_path / "trunc.stz")
    with open(p, "wb") as f:
        f.write(gzip.compress(good * 50)[:-20])       # compressed stream cut short
    with pytest.raises(H.GoChemError):
        H.stf_read(p, 100, 2)


def _lzw_msb_decode(data: bytes) -> bytes:
    """Go's compress/lzw reader with lzw.MSB and 8-bit literals, restated: 9..12-bit codes packed most significant bit first,
    256 = clear, 257 = end of data, entries from 258, the width grows when the next free entry reaches 1 << width."""
    out = bytearray()
    prefix, suffix = {}, {}
    width, hi, overflow, last = 9, 257, 512, None
    bits = nbits = 0
    pos = 0
    while True:
        while nbits < width:
            bits = (bits << 8) | data[pos]
            pos += 1
            nbits += 8
        code = (bits >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        if code == 256:
            width, hi, overflow, last = 9, 257, 512, None
            continue
        if code == 257:
            return bytes(out)
        if code < 256:
            exp = bytes([code])
        elif code <= hi:
            def expand(c):
                b = bytearray()
                while c >= 256:
                    b.append(suffix[c])
                    c = prefix[c]
                b.append(c)
                return bytes(reversed(b))
            if code == hi and last is not None:
                e = expand(last)
                exp = e + e[:1]
            else:
                exp = expand(code)
        else:
            raise ValueError("lzw: invalid code")
        if last is not None:
            suffix[hi], prefix[hi] = exp[0], last
        out += exp
        last, hi = code, hi + 1
        if hi >= overflow:
            if width == 12:
                last, hi = None, hi - 1
            else:
                width += 1
                overflow = 1 << width


def test_stf_lzw_form(tmp_path):
    """The 'l' files (lzw.NewWriter / NewReader with lzw.MSB and 8-bit literals, stf.go:143, 262): what StfW writes decodes with an
    independent restatement of Go's reader to the expected text, and StfR reads it back -- a stream long enough for the code width to
    grow to 12 bits and the dictionary to be dropped several times."""
    rng = np.random.default_rng(9)
    frames = rng.normal(0, 25, (40, 400, 3))
    want = np.rint(frames * 100.0) / 100.0
    p = str(tmp_path / "w.stl")
    H.stf_write(p, frames, header={"codec": "lzw"})
    raw = open(p, "rb").read()
    text = _stf_text(frames, header={"codec": "lzw"})
    assert len(text) > 200_000 and len(raw) < len(text)
    assert _lzw_msb_decode(raw) == text
    got, nf, na, header = H.stf_read(p, 50, 400)
    assert (nf, na, header) == (40, 400, {"codec": "lzw"}) and np.array_equal(got, want)
    raw_feed, nf2, _, _ = H.stf_read(p, 50, 400, raw=True)
    assert nf2 == 40 and np.array_equal(raw_feed, want)
    # a tiny stream (no width change), and a truncated one
    H.stf_write(str(tmp_path / "tiny.stl"), frames[:1, :3])
    got, nf, na, _ = H.stf_read(str(tmp_path / "tiny.stl"), 2, 3)
    assert (nf, na) == (1, 3) and np.array_equal(got[0], want[0, :3])
    with open(str(tmp_path / "cut.stl"), "wb") as f:
        f.write(raw[: len(raw) // 2])
    with pytest.raises(H.GoChemError):
        H.stf_read(str(tmp_path / "cut.stl"), 50, 400)


def test_rhos_and_shape_indexes():
    """chem.Rhos (geometric.go:523-545): eigenvalues of the moment tensor, largest first; chem.RhoShapeIndexes (:511-520);
    the 'collapsed to a single point' error when even the largest eigenvalue vanishes."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = rng.normal(0, 1, (3, 3)) * rng.uniform(0.1, 30)
        m = a @ a.T
        rhos, lin, circ = H.rhos_shape(m)
        want = np.sort(np.linalg.eigvalsh(m))[::-1]
        assert np.abs(rhos - want).max() < 1e-10 * max(1.0, want[0])
        assert abs(lin - (1 - want[1] / want[0]) * 100) < 1e-8 and abs(circ - (1 - want[2] / want[1]) * 100) < 1e-6
    rhos, lin, circ = H.rhos_shape(np.diag([4.0, 1.0, 0.25]))
    assert rhos.tolist() == [4.0, 1.0, 0.25] and lin == 75.0 and circ == 75.0
    with pytest.raises(H.GoChemError, match="colapsed to a single point"):
        H.rhos_shape(np.zeros((3, 3)))
