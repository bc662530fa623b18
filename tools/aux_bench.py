#!/usr/bin/env python
"""Logging rollouts (state_dot + 18 outputs stored every k-th step) on config 2 (development tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

eng = mp.Engine(0)
p = mp.VehicleParameters()
p.DFL = p.DFR = p.DRL = p.DRR = 1.0
eng.set_params(p)
s0, d, t = wl.config2_rollouts()
a, b, c = eng.dev(s0), eng.dev(d), eng.dev(t)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in ("auto", "closed_form"):
    eng.set_friction_mode(mode)
    for stride in (1, 10, 50):
        for k in range(3):
            if k == 2:
                e0.record()
            r = eng.rollout(a, b, c, wl.DT, 500, hold=wl.HOLD, store_stride=stride, want_aux=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"{mode:12s} stride {stride:3d}: {e0.elapsed_time(e1):.3f} ms  {65536 * 500 / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
        del r
