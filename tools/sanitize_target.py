#!/usr/bin/env python
"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck): tiny sizes, odd shapes."""
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
s0, d, t = wl.config2_rollouts(B=333, n_steps=40)
for mode in ("auto", "closed_form"):
    eng.set_friction_mode(mode)
    eng.rollout(s0, d, t, wl.DT, 40, hold=10, store_stride=1)
    eng.rollout(s0, d, t, wl.DT, 40, hold=10, store_stride=5, want_aux=True)
    eng.rollout(s0, d, t, wl.DT, 40, hold=10, dtype="f32", store_stride=2)
    st0, wps = wl.tracking_fleet(V=200, n_sets=3, W=777)
    eng.track_closed_loop(st0, wps, wl.DT, 37, 25.0, vehicles_per_set=67, store_stride=1, want_log=True, want_target_idx=True)
    eng.track_closed_loop(st0, wps, wl.DT, 37, 25.0, vehicles_per_set=67)
eng.set_friction_mode("auto")
# sliced launch (more than one wave of CTAs)
s0, d, t = wl.config2_rollouts(B=148 * 64 * 5, n_steps=60)
eng.rollout(s0, d, t, wl.DT, 60, hold=10, store_stride=0)
w = wl.config3_lattice(P=77, M=2500)
for mode in ("auto", "fp64"):
    eng.set_collision_mode(mode)
    f = eng.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"])
eng.set_collision_mode("auto")
eng.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"], want_clearance=True)
eng.select_best_path_index_batch(w["px"][:, -1].copy(), w["py"][:, -1].copy(), f, w["goal"], w["weight"])
par = w["spiral_params"]
lat = eng.sample_lattice(par[0], par[1], par[2], ego=w["ego"])
o = eng.optimize_spirals(np.linspace(20, 40, 77), np.linspace(-5, 5, 77), np.linspace(-0.3, 0.3, 77))
eng.plan_lattice(np.stack([np.linspace(20, 40, 21), np.linspace(-5, 5, 21), np.linspace(-0.3, 0.3, 21)]), (1.0, 2.0, 0.1),
                 w["obstacles"], w["offsets"], w["radii"], (40.0, 5.0), 10.0)
dd, tt = eng.mpc_sample_controls(1000, 10, 7)
eng.argmin(torch.rand(5000, dtype=torch.float64, device=eng.tdev))
torch.cuda.synchronize()
print("sanitize target done")
