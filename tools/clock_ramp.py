"""How long must the GPU be kept busy after idling before a short, issue-bound kernel is timed at its boost clock, and when does
the power cap arrive?  Idle 8 s, run the per-atom pass for `warm` seconds, then time 3 x 5 passes.  python tools/clock_ramp.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pynvml

from gochem_b200 import _lib, synth

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
natoms, nframes = 5000, 50000
ref, sig = synth.make_reference(natoms), synth.atom_sigmas(natoms)
t = _lib.DeviceTraj(natoms, nframes)
t.synth_fill(ref, sig, nframes)
idx = np.arange(0, natoms, 2, dtype=np.int32)
t.peratom_dev_sum(ref, idx)
for warm in (0.0, 0.02, 0.05, 0.1, 0.2, 0.4, 1.0, 3.0):
    time.sleep(8.0)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < warm:
        t.peratom_dev_sum(ref, idx)
    rounds, clocks = [], []
    for _ in range(3):
        t.timer_start()
        for _ in range(5):
            t.peratom_dev_sum(ref, idx)
        rounds.append(t.timer_stop() / 5)
        clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    print(json.dumps({"warm_s": warm, "pass_ms_rounds": [round(r, 4) for r in rounds], "sm_mhz_after_each_round": clocks,
                      "power_w": pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0}), flush=True)
