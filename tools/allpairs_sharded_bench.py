"""BASELINE configs[3] on N GPUs of one node: all-pairs RMSD matrix of F frames x N atoms with the tiles dealt round-robin and every
finished tile stored straight into its owners' row blocks over NVLink (gcb_allpairs_rmsd_sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/allpairs_sharded_bench.py [nframes] [natoms] [fit]

Frames start sharded (each rank synthesises and superposes its own block), are all-gathered with NCCL, then the matrix is computed.
Times are device times (CUDA events on the library's stream), max over ranks."""
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

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
natoms = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
fit = len(sys.argv) > 3 and sys.argv[3] == "fit"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
comm = _lib.Comm(local, world, rank, bytes(uid.cpu().tolist()))
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
counts = [H.shard_range(nframes, r, world)[1] for r in range(world)]
first = sum(counts[:rank])
t = _lib.DeviceTraj(natoms, nframes, dev=local)
t.synth_fill(ref, sig, counts[rank], first=first, frame_id0=first)
t.set_nframes(nframes)
t.super_rmsd(ref, first=first, nframes=counts[rank], write_back=True, want=())    # definition (i): frames pre-superposed
t.sync()
dist.barrier()
t.allgather(comm, counts)        # first NCCL broadcast: connection set-up
dist.barrier()
t.timer_start()
t.allgather(comm, counts)
gather_ms = t.timer_stop()
_, row0, gms, tms, redone = t.allpairs_rmsd_sharded(comm, fit=fit, fetch=False)       # warm-up: blocks allocated and mapped
reps = 3
res = []
for _ in range(reps):
    _, row0, gms, tms, redone = t.allpairs_rmsd_sharded(comm, fit=fit, fetch=False)
    res.append((gms, tms))
blk, row0, gms, tms, redone = t.allpairs_rmsd_sharded(comm, fit=fit, fetch=True)   # first touch of the host block
t0 = time.perf_counter()
blk, row0, gms, tms, redone = t.allpairs_rmsd_sharded(comm, fit=fit, fetch=True)
d2h_s = time.perf_counter() - t0 - tms * 1e-3
res.append((gms, tms))
g = [None] * world
dist.all_gather_object(g, {"rank": rank, "rows": [row0, blk.shape[0]], "gemm_ms": [r[0] for r in res], "total_ms": [r[1] for r in res],
                           "allgather_ms": gather_ms, "d2h_s": d2h_s, "mean": float(blk.mean()), "redone": redone,
                           "diag_zero": bool((blk[np.arange(blk.shape[0]), row0 + np.arange(blk.shape[0])] == 0).all())})
if rank == 0:
    total = max(min(x["total_ms"]) for x in g)
    gemm = max(min(x["gemm_ms"]) for x in g)
    npad = (natoms + 15) // 16 * 16
    tf, side, kdim = (32, 96, npad) if fit else (128, 128, 3 * npad)
    nt = (nframes + tf - 1) // tf
    flops = 2.0 * (nt * (nt + 1) // 2) * side * side * kdim
    print(json.dumps({"variant": "per-pair fit" if fit else "no fit", "nframes": nframes, "natoms": natoms, "world": world,
                      "total_ms_max_over_ranks": total, "gemm_ms_max_over_ranks": gemm, "frames_allgather_ms": max(x["allgather_ms"] for x in g),
                      "dmma_tflops_aggregate": flops / (total * 1e-3) / 1e12, "d2h_rows_s": max(x["d2h_s"] for x in g),
                      "matrix_mean_rmsd": float(np.average([x["mean"] for x in g], weights=[x["rows"][1] for x in g])),
                      "per_rank": g}), flush=True)
dist.barrier()
t.close()
comm.close()
dist.destroy_process_group()
