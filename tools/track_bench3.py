#!/usr/bin/env python
"""Closed-loop kernel: marginal cost of a control update (development tool): N and ctrl_every swept at 65,536 vehicles."""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
st0, wps = wl.tracking_fleet(V=B, n_sets=16)
s, w = eng.dev(st0), eng.dev(wps)
for N, ce in ((100, 10), (500, 10), (1000, 10), (500, 5), (500, 20), (500, 50), (500, 250), (500, 500)):
    for k in range(4):
        if k == 2:
            e0.record()
        r = eng.track_closed_loop(s, w, 1e-4, N, 25.0, vehicles_per_set=-(-B // 16), ctrl_every=ce)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"track B={B} N={N} ctrl_every={ce}: {ms:.3f} ms  {B * N / ms * 1e3:.3e} steps/s  ({N // ce} updates)")
for N in (100, 500, 1000):
    d = torch.zeros(N // 10, 1, B, dtype=torch.float64, device=eng.tdev)
    for k in range(4):
        if k == 2:
            e0.record()
        eng.rollout(s, d, d, 1e-4, N, hold=10, store_stride=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"open  B={B} N={N}: {ms:.3f} ms  {B * N / ms * 1e3:.3e} steps/s")
