"""BASELINE configs[3]: all-pairs RMSD matrix, F frames x N atoms, on one GPU (or one row block per rank).
    python tools/allpairs_bench.py [nframes] [natoms]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gochem_b200 import _lib, synth

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
natoms = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
fit = len(sys.argv) > 3 and sys.argv[3] == "fit"
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = int(os.environ.get("LOCAL_RANK", "0"))
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes, dev=dev)
t.synth_fill(ref, sig, nframes)
t.super_rmsd(ref, write_back=True, want=())          # definition (i): frames pre-superposed on the template
base, rem = divmod(nframes, world)
row0 = rank * base + min(rank, rem)
nrows = base + (1 if rank < rem else 0)
out, ms, redone = t.allpairs_rmsd(row0=row0, nrows=nrows, fit=fit)   # warm-up (allocations, attribute calls)
t0 = time.perf_counter()
out, ms, redone = t.allpairs_rmsd(row0=row0, nrows=nrows, fit=fit)
wall = time.perf_counter() - t0
npad = (natoms + 15) // 16 * 16
tf = 32 if fit else 128                      # frames per tile side
kdim = npad if fit else 3 * npad
side = 96 if fit else 128                    # GEMM rows per tile side
tiles = (nframes + tf - 1) // tf
ntiles = sum(1 for bi in range(tiles) for bj in range(bi, tiles)
             if (bi * tf < row0 + nrows and bi * tf + tf > row0) or (bj * tf < row0 + nrows and bj * tf + tf > row0))
flops_done = 2.0 * ntiles * side * side * kdim
res = {"variant": "per-pair fit" if fit else "no fit", "nframes": nframes, "natoms": natoms, "rank": rank, "world": world, "rows": [row0, nrows], "gemm_ms": ms,
       "wall_s_incl_d2h": wall, "tiles": ntiles, "dmma_tflops": flops_done / (ms * 1e-3) / 1e12,
       "useful_tflops_full_matrix_equiv": 2.0 * nrows * nframes * 3 * natoms / (ms * 1e-3) / 1e12, "pairs_recomputed": redone,
       "mean_rmsd": float(out.mean())}
# spot check against the single-pair entry point on the same resident frames
fr = t.download(0, 3)
res["spot_check_abs_err"] = float(abs(out[0, 1] - np.sqrt(((fr[0] - fr[1]) ** 2).sum() / natoms))) if (row0 == 0 and not fit) else None
if rank == 0 and os.environ.get("CUBLAS_REF", "1") == "1":
    import torch

    n = min(nframes, 16384)
    a = torch.randn(n, kdim, dtype=torch.float64, device=f"cuda:{dev}")
    torch.matmul(a, a.T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, a.T)
    e1.record()
    torch.cuda.synchronize()
    res["cublas_dgemm_tflops"] = 2.0 * n * n * kdim / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps(res), flush=True)
