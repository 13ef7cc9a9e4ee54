"""BASELINE configs[2]: align.LOVO on a resident trajectory, frames sharded over the ranks (1 rank without torchrun).
    python tools/lovo_bench.py [natoms] [nframes_total]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/lovo_bench.py
Reports the whole LOVO run (iterations, wall seconds, frames/s per iteration = total frames x iterations / seconds) after one
warm-up run, and the wall time of a single RMSDTraj pass (kernel + all-reduce + sums back on the host) beside its kernel time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
total = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
comm = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    comm = _lib.Comm(local, world, rank, bytes(uid.cpu().tolist()))
    barrier = dist.barrier
else:
    barrier = lambda: None
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
first, count = H.shard_range(total, rank, world)
t = _lib.DeviceTraj(natoms, count, dev=local)
t.synth_fill(ref, sig, count, frame_id0=first)
t.sync()
kw = dict(nselected_total=total, cpus=8, comm=comm, rank=rank, nranks=world)
H.LOVOResident(t, ref, **kw)                      # warm-up: allocations, NCCL connections
runs = []
for _ in range(3):
    barrier()
    t0 = time.perf_counter()
    got = H.LOVOResident(t, ref, **kw)
    t.sync()
    barrier()
    runs.append(time.perf_counter() - t0)
best = min(runs)
# one pass: wall (kernel + reduce + all-reduce + D2H) vs device time of the pass
idx = np.arange(natoms)
passes = []
for _ in range(5):
    barrier()
    t0 = time.perf_counter()
    t.peratom_dev_sum(ref, idx, comm=comm)
    passes.append(time.perf_counter() - t0)
t.timer_start()
for _ in range(5):
    t.peratom_dev_sum(ref, idx, comm=comm)
dev_ms = t.timer_stop() / 5
if rank == 0:
    print(json.dumps({"natoms": natoms, "frames_total": total, "world": world, "N": got["N"], "iterations": got["Iterations"],
                      "lovo_seconds_best_of_3": best, "lovo_seconds_runs": runs,
                      "frames_per_s_per_iteration": total * got["Iterations"] / best,
                      "one_pass_wall_ms_min": min(passes) * 1e3, "one_pass_stream_ms": dev_ms,
                      "one_pass_frames_per_s_wall": total / min(passes)}), flush=True)
barrier()
t.close()
if comm is not None:
    comm.close()
    dist.destroy_process_group()
