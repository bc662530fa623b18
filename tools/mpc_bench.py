#!/usr/bin/env python
"""Config 4 on one GPU: time split of control sampling / rollout with running cost / argmin (development tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

eng = mp.Engine(0)
p = mp.VehicleParameters()
eng.set_params(p)
cfg = wl.config4_mpc(B=1 << 20)
B, N = cfg["B"], cfg["n_steps"]
s0 = eng.dev(cfg["state0"]).reshape(12, 1).expand(12, B).contiguous()
cref = eng.dev(cfg["cost_ref"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for k in range(4):
    ev[0].record()
    d, t = eng.mpc_sample_controls(B, N, cfg["seed"])
    ev[1].record()
    r = eng.rollout(s0, d, t, wl.DT, N, hold=1, cost_ref=cref, w_u=cfg["w_u"], u_ref=cfg["u_ref"])
    ev[2].record()
    mn, ix = eng.argmin(r.cost)
    ev[3].record()
torch.cuda.synchronize()
print(f"sample {ev[0].elapsed_time(ev[1]):.3f} ms  rollout+cost {ev[1].elapsed_time(ev[2]):.3f} ms  argmin {ev[2].elapsed_time(ev[3]):.3f} ms  "
      f"total {ev[0].elapsed_time(ev[3]):.3f} ms = {B * N / ev[0].elapsed_time(ev[3]) * 1e3:.3e} steps/s; best {int(ix.item())} {float(mn.item()):.6e}")
for hold in (1, 10):
    for k in range(3):
        ev[0].record()
        r = eng.rollout(s0, d, t, wl.DT, N, hold=hold)
        ev[1].record()
    torch.cuda.synchronize()
    print(f"plain rollout hold={hold}: {ev[0].elapsed_time(ev[1]):.3f} ms")
