#!/usr/bin/env python
"""FP32 twin of the rollout kernel: timing and drift against FP64 (development tool)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

eng = mp.Engine(0)
p = mp.VehicleParameters()
p.DFL = p.DFR = p.DRL = p.DRR = 1.0
eng.set_params(p)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0, d, t = wl.config2_rollouts()
for B, N, stride in ((65536, 500, 1), (65536, 500, 0), (1048576, 100, 0)):
    if B != 65536:
        s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    for dt in ("f64", "f32"):
        td = torch.float64 if dt == "f64" else torch.float32
        a, b, c = eng.dev(s0, td), eng.dev(d, td), eng.dev(t, td)
        for k in range(4):
            if k == 3:
                e0.record()
            r = eng.rollout(a, b, c, wl.DT, N, hold=wl.HOLD, store_stride=stride, dtype=dt)
        e1.record()
        torch.cuda.synchronize()
        print(f"{dt} B={B} N={N} stride={stride}: {e0.elapsed_time(e1):.3f} ms  {B * N / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
        if dt == "f64":
            ref = r.state_end.cpu().numpy()
        else:
            e = np.abs(r.state_end.cpu().numpy().astype(np.float64) - ref) / np.maximum(np.abs(ref), 1.0)
            print("   drift f32 vs f64 after", N, "steps: max", e.max(), "per component", e.max(axis=1).round(8))
        del r
