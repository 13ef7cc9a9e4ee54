"""solv.ConcMolRDF's frame loop on resident frames: python tools/rdf_bench.py [nwater] [nsolute] [nframes]
A solute of nsolute atoms in a cubic box of nwater 3-atom waters at liquid density; CPU oracle timed on a few frames beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from gochem_b200 import _lib

nwater = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
nsolute = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
nframes = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
rng = np.random.default_rng(7)
box = (nwater / 0.0334) ** (1.0 / 3.0)          # 0.0334 waters per cubic Angstrom
sol = rng.normal(0, (nsolute / 1500.0) ** (1.0 / 3.0) * 9.0, (nsolute, 3))
wo = rng.uniform(-box / 2, box / 2, (nwater, 3))
wat = (wo[:, None, :] + np.concatenate([np.zeros((nwater, 1, 3)), rng.normal(0, 0.6, (nwater, 2, 3))], axis=1)).reshape(-1, 3)
base = np.concatenate([sol, wat])
natoms = base.shape[0]
molid = np.concatenate([np.repeat(np.arange(1, nsolute // 10 + 2), 10)[:nsolute], nsolute + np.repeat(np.arange(nwater), 3)]).astype(np.int32)
chain = np.concatenate([np.ones(nsolute), np.full(3 * nwater, 2)]).astype(np.int32)
flags = np.concatenate([np.zeros(nsolute), np.full(3 * nwater, 3)]).astype(np.uint8)
refidx = np.arange(nsolute)
sample = 64
frames = (base[None] + rng.normal(0, 0.5, (sample, natoms, 3))).astype(np.float32).astype(np.float64)
t = _lib.DeviceTraj(natoms, nframes)
for f0 in range(0, nframes, sample):
    n = min(sample, nframes - f0)
    t.upload(frames[:n], first=f0)
t.rdf_cdf_sum(molid, chain, flags, refidx)
t.sync()
t.timer_start()
got, read = t.rdf_cdf_sum(molid, chain, flags, refidx)
ms = t.timer_stop()
ncpu = 8
t0 = time.perf_counter()
want, _ = oracle.rdf_cdf_sum(frames[:ncpu], molid, chain, flags, refidx)
cpu_s = (time.perf_counter() - t0) / ncpu
chk, _ = t.rdf_cdf_sum(molid, chain, flags, refidx, nframes=ncpu)
print(json.dumps({"natoms": natoms, "nsolute": nsolute, "nwater": nwater, "nframes": nframes, "gpu_ms": ms, "frames_per_s": nframes / ms * 1e3,
                  "cpu_oracle_frames_per_s_1thread": 1.0 / cpu_s, "speedup_vs_1_cpu_thread": (nframes / ms * 1e3) * cpu_s,
                  "counts_equal_oracle_first_%d_frames" % ncpu: bool(np.array_equal(chk, want)), "last_shell_mean_count": float(got[-1] / read)}))
