#!/usr/bin/env python
"""Closed-loop tracking kernel timing (development tool): fleets of several sizes vs the open-loop rollout."""
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
for B, N in ((65536, 500), (75776, 500), (37888, 500), (303104, 200), (1048576, 100)):
    st0, wps = wl.tracking_fleet(V=B, n_sets=16)
    s, w = eng.dev(st0), eng.dev(wps)
    for kw in ({}, {"store_stride": 10, "want_log": True}):
        for k in range(3):
            if k == 2:
                e0.record()
            r = eng.track_closed_loop(s, w, 1e-4, N, 25.0, vehicles_per_set=-(-B // 16), **kw)
        e1.record()
        torch.cuda.synchronize()
        print(f"track B={B} N={N} {kw} {e0.elapsed_time(e1):.3f} ms  {B * N / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
        del r
    d = torch.zeros(N // 10, 1, B, dtype=torch.float64, device=eng.tdev)
    for k in range(3):
        if k == 2:
            e0.record()
        eng.rollout(s, d, d, 1e-4, N, hold=10, store_stride=0)
    e1.record()
    torch.cuda.synchronize()
    print(f"open  B={B} N={N} {e0.elapsed_time(e1):.3f} ms  {B * N / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
