"""BASELINE configs[4] through the reader: one DCD file, all ranks of a node stream it at once (align.RMSDTrajStreamed, windows dealt to the
ranks in turn, every rank reads only its own windows), per-atom sums combined with NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/dcd_stream_multi.py [natoms] [nframes]"""
import json, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from gochem_b200 import _lib, synth
from gochem_b200 import host_api as H

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
comm = _lib.Comm(local, world, rank, bytes(uid.cpu().tolist()))
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
path = os.path.join(tempfile.gettempdir(), "gcb_stream_multi.dcd")
if rank == 0:
    t = _lib.DeviceTraj(natoms, nframes, dev=local)
    t.synth_fill(ref, sig, nframes, through_f32=True)
    H.dcd_write(path, t.download())
    t.close()
dist.barrier()
size = os.path.getsize(path)
idx = np.arange(natoms)
kw = dict(dev=local, rank=rank, nranks=world, comm=comm.h)
H.RMSDTrajStreamed(path, ref, idx, 8, 0, **kw)           # warm-up
times = []
for _ in range(3):
    dist.barrier()
    t0 = time.perf_counter()
    got, n = H.RMSDTrajStreamed(path, ref, idx, 8, 0, **kw)
    dist.barrier()
    times.append(time.perf_counter() - t0)
g = [None] * world
dist.all_gather_object(g, got.tobytes())
if rank == 0:
    single, n1 = H.RMSDTrajStreamed(path, ref, idx, 8, 0, dev=local)
    dt = min(times)
    print(json.dumps({"natoms": natoms, "nframes": n, "world": world, "file_GB": size / 1e9, "seconds_best_of_3": dt, "runs": times,
                      "frames_per_s": n / dt, "file_GBs": size / dt / 1e9, "identical_across_ranks": all(x == g[0] for x in g),
                      "max_abs_diff_vs_one_rank": float(np.abs(got - single).max())}), flush=True)
dist.barrier()
if rank == 0:
    os.remove(path)
comm.close()
dist.destroy_process_group()
