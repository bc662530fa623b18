#!/usr/bin/env python
"""Latency of the scalar drop-in calls (one vehicle, one step per call): config 1's inner loop (development tool)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402

vm = mp.VehicleModel(dt=1e-4)
p = mp.VehicleParameters()
st = [20.0, 0.0, 0.0] + [20.0 / p.rw] * 4 + [0.0, 0.0, 0.0]
ax = ay = 0.0
for name in ("planar_model_RK4", "planar_model"):
    fn = getattr(vm, name)
    for k in range(1030):
        if k == 30:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        r = fn(st, [50.0] * 4, [1.0] * 4, [0.01, 0.01, 0.0, 0.0], p, ax, ay)
        if name == "planar_model_RK4":
            st, ax, ay = list(r[0]), r[7], r[8]
    print(f"{name}: {(time.perf_counter() - t0) / 1000 * 1e6:.1f} us per call")
