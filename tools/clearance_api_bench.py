#!/usr/bin/env python
"""Wall time of Engine.collision_check_batch(..., want_clearance=True) from numpy inputs, the three clearance_trig modes (config 3)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

eng = mp.Engine(0)
w = wl.config3_lattice()
px, py, obs = eng.dev(w["px"]), eng.dev(w["py"]), eng.dev(w["obstacles"])
for m in ("auto", "host", "device"):
    for rep in range(8):
        if rep == 3:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        f, c = eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"], want_clearance=True, clearance_trig=m)
        c_host = c.cpu()
    dt = (time.perf_counter() - t0) / 5
    print(f"clearance_trig={m}: {dt * 1e3:.3f} ms per call (flags + clearance on the host), candidates {getattr(eng, 'last_clearance_candidates', None)}")
