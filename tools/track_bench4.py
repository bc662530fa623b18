#!/usr/bin/env python
"""Closed-loop kernel, short A/B timing (development tool)."""
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
B = 65536
st0, wps = wl.tracking_fleet(V=B, n_sets=16)
s, w = eng.dev(st0), eng.dev(wps)
out = []
for N, ce, kw in ((500, 10, {}), (1000, 10, {}), (500, 20, {}), (500, 5, {}), (500, 10, {"store_stride": 10, "want_log": True})):
    for k in range(5):
        if k == 2:
            e0.record()
        r = eng.track_closed_loop(s, w, 1e-4, N, 25.0, vehicles_per_set=-(-B // 16), ctrl_every=ce, **kw)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"N={N}/ce={ce}{'/log' if kw else ''}: {e0.elapsed_time(e1) / 3:.3f} ms")
print(sys.argv[1] if len(sys.argv) > 1 else "", " | ".join(out), " checksum", float(r.state_end.sum()))
