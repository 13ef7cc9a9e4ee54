"""Run under torchrun on >= 2 GPUs: NCCL rank-ordered allreduce and sharded LOVO against the single-GPU answer.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# the library's own communicator: rank 0 makes the id, torch.distributed carries the 128 bytes
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
comm = _lib.Comm(local, world, rank, bytes(uid.cpu().tolist()))
out = {}
# 1. rank-ordered sum is bit-identical on every rank
v = np.random.default_rng(rank).standard_normal(5000)
s = comm.allreduce_sum(v)
want = sum(np.random.default_rng(r).standard_normal(5000) for r in range(world))
out["allreduce_max_abs_err"] = float(np.abs(s - want).max())
g = [None] * world
dist.all_gather_object(g, s.tobytes())
out["allreduce_bit_identical"] = all(x == g[0] for x in g)
# 2. sharded LOVO (config 3 shape) vs the single-GPU run
natoms = int(os.environ.get("NATOMS", "5000"))
total = int(os.environ.get("NFRAMES", "50000"))
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
first, count = H.shard_range(total, rank, world)
t = _lib.DeviceTraj(natoms, count, dev=local)
t.synth_fill(ref, sig, count, frame_id0=first)
t.sync()
dist.barrier()
t0 = time.perf_counter()
got = H.LOVOResident(t, ref, nselected_total=total, cpus=8, comm=comm, rank=rank, nranks=world)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
out["lovo"] = {"N": got["N"], "iterations": got["Iterations"], "seconds": dt,
               "frames_per_s_per_iteration": total * got["Iterations"] / dt}
g = [None] * world
dist.all_gather_object(g, got["Natoms"].tobytes())
out["natoms_identical_across_ranks"] = all(x == g[0] for x in g)
# 3. all-pairs matrix: frames sharded -> all-gather over NVLink -> tiles dealt round-robin, stored straight into the owners' rows
ap_frames, ap_atoms = int(os.environ.get("AP_FRAMES", "3000")), int(os.environ.get("AP_ATOMS", "600"))
aref, asig = synth.make_reference(ap_atoms), synth.atom_sigmas(ap_atoms)
counts = [H.shard_range(ap_frames, r, world)[1] for r in range(world)]
afirst = sum(counts[:rank])
ta = _lib.DeviceTraj(ap_atoms, ap_frames, dev=local)
ta.synth_fill(aref, asig, counts[rank], first=afirst, frame_id0=afirst)      # only this rank's shard is filled
ta.set_nframes(ap_frames)
ta.super_rmsd(aref, first=afirst, nframes=counts[rank], write_back=True, want=())
ta.allgather(comm, counts)
for fit in (False, True):
    blk, row0, gms, tms, redone = ta.allpairs_rmsd_sharded(comm, fit=fit)
    blk2, _, gms, tms, redone = ta.allpairs_rmsd_sharded(comm, fit=fit)
    want, _, _ = ta.allpairs_rmsd(row0=row0, nrows=blk.shape[0], fit=fit)   # same rows computed locally by the single-GPU path
    ok = bool(np.array_equal(blk, want) and np.array_equal(blk2, want))
    g = [None] * world
    dist.all_gather_object(g, (ok, gms, tms))
    out["allpairs_fit" if fit else "allpairs"] = {"blocks_equal_single_gpu": all(x[0] for x in g), "gemm_ms": [x[1] for x in g],
                                                  "total_ms": [x[2] for x in g], "frames": ap_frames, "atoms": ap_atoms}
ta.close()
# 4. solvent shells (gcb_rdf_cdf_sum): frames sharded, integer counts summed over the ranks == single-GPU counts
rng = np.random.default_rng(5)
nsol, nwat, rf = 60, 900, 64 * world
base = np.concatenate([rng.normal(0, 4, (nsol, 3)), np.repeat(rng.uniform(-20, 20, (nwat, 3)), 3, axis=0) + rng.normal(0, 0.5, (3 * nwat, 3))])
rframes = base[None] + rng.normal(0, 0.4, (rf, base.shape[0], 3))
molid = np.concatenate([np.repeat(np.arange(1, 7), 10), 7 + np.repeat(np.arange(nwat), 3)]).astype(np.int32)
chain = np.concatenate([np.ones(nsol), np.full(3 * nwat, 2)]).astype(np.int32)
flags = np.concatenate([np.zeros(nsol), np.full(3 * nwat, 3)]).astype(np.uint8)
r0, rc = H.shard_range(rf, rank, world)
tr = _lib.DeviceTraj(base.shape[0], max(rc, 1), dev=local)
tr.upload(rframes[r0:r0 + rc])
got_cdf, got_read = tr.rdf_cdf_sum(molid, chain, flags, np.arange(nsol), comm=comm)
tr.close()
g = [None] * world
dist.all_gather_object(g, got_cdf.tobytes())
out["rdf"] = {"identical_across_ranks": all(x == g[0] for x in g), "frames_read": got_read}
if rank == 0:
    tw = _lib.DeviceTraj(base.shape[0], rf, dev=local)
    tw.upload(rframes)
    want_cdf, want_read = tw.rdf_cdf_sum(molid, chain, flags, np.arange(nsol))
    tw.close()
    out["rdf"]["equal_single_gpu"] = bool(np.array_equal(want_cdf, got_cdf) and want_read == got_read)
if rank == 0:
    whole = _lib.DeviceTraj(natoms, total, dev=local)
    whole.synth_fill(ref, sig, total)
    t0 = time.perf_counter()
    single = H.LOVOResident(whole, ref, cpus=8)
    out["single_gpu_seconds"] = time.perf_counter() - t0
    out["same_as_single_gpu"] = bool(single["N"] == got["N"] and list(single["Natoms"]) == list(got["Natoms"])
                                     and single["Iterations"] == got["Iterations"])
    out["max_abs_rmsd_diff_vs_single"] = float(np.abs(single["RMSD"] - got["RMSD"]).max())
    whole.close()
# 5. last (the communicator is unusable afterwards): a peer that never delivers its vector gives an error after
#    GCB_EXCHANGE_TIMEOUT_S seconds, not a hung GPU.  Only run when that variable is set (to something short).
if os.environ.get("GCB_EXCHANGE_TIMEOUT_S"):
    if rank == 0:
        t0 = time.perf_counter()
        try:
            t.peratom_dev_sum(ref, np.arange(natoms, dtype=np.int32), comm=comm)
            out["exchange_timeout"] = "no error raised"
        except _lib.GcbError as e:
            out["exchange_timeout"] = {"seconds": time.perf_counter() - t0, "error": e.msg[:160]}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
t.close()
comm.close()
dist.destroy_process_group()
