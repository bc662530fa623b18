#!/usr/bin/env python
"""Short single-GPU program for ncu captures: a few launches of one hot-path kernel at full size.

    python tools/prof_target.py rollout|rollout_f32|collision|collision_clear|planner|mpc|track|track_log [--launches 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--steps", type=int, default=100, help="track / track_log: steps per launch")
    a = ap.parse_args()
    eng = mp.Engine(0)
    p = mp.VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    eng.set_params(p)
    if a.what in ("rollout", "rollout_f32"):
        dt = "f32" if a.what.endswith("f32") else "f64"
        td = torch.float32 if dt == "f32" else torch.float64
        s0, d, t = wl.config2_rollouts()
        s0, d, t = eng.dev(s0, td), eng.dev(d, td), eng.dev(t, td)
        traj = eng.empty(500, 10, 65536, dtype=td)
        for _ in range(a.launches):
            eng.rollout(s0, d, t, wl.DT, 500, hold=wl.HOLD, store_stride=1, traj_out=traj, dtype=dt)
    elif a.what in ("collision", "collision_clear"):
        w = wl.config3_lattice()
        for _ in range(a.launches):
            eng.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"],
                                      want_clearance=a.what.endswith("clear"))
    elif a.what in ("track", "track_log"):
        st0, wps = wl.tracking_fleet(V=65536, n_sets=16)
        s, w = eng.dev(st0), eng.dev(wps)
        for _ in range(a.launches):
            eng.track_closed_loop(s, w, wl.DT, a.steps, 25.0, vehicles_per_set=4096,
                                  **({"store_stride": 10, "want_log": True} if a.what.endswith("log") else {}))
    elif a.what == "planner":
        # one launch of every planner-side kernel at config-3 size: obstacle_prepare + collision_cull<3> (flags from yaws),
        # obstacle_prepare + clearance_cull<3> (+ reduce), select_score + argmin, lattice_kernel, spiral_opt_kernel
        import numpy as np
        w = wl.config3_lattice()
        px, py, yaw, obs = eng.dev(w["px"]), eng.dev(w["py"]), eng.dev(w["pyaw"]), eng.dev(w["obstacles"])
        trig = eng.path_trig(w["pyaw"], w["px"].shape[1])
        par = w["spiral_params"]
        k1d, k2d, sfd = eng.dev(par[0]), eng.dev(par[1]), eng.dev(par[2])
        lat0 = eng.sample_lattice(k1d, k2d, sfd, ego=None, want_trig=False)
        tf_goal = eng.dev(3.0 * (par[0] + par[1]) * par[2] / 8.0)
        for _ in range(a.launches):
            free = eng.collision_check_batch(px, py, yaw, obs, w["offsets"], w["radii"])
            eng.collision_check_batch(px, py, None, obs, w["offsets"], w["radii"], trig=trig, want_clearance=True)
            eng.select_best_path_index_batch(px[:, -1].contiguous(), py[:, -1].contiguous(), free, w["goal"], w["weight"])
            eng.sample_lattice(k1d, k2d, sfd, ego=eng.dev(w["ego"]))
            eng.optimize_spirals(lat0["end_xy"][0], lat0["end_xy"][1], tf_goal)
    elif a.what == "mpc":
        cfg = wl.config4_mpc(B=1 << 20)
        d, t = eng.mpc_sample_controls(cfg["B"], 100, cfg["seed"])
        s0 = eng.dev(cfg["state0"]).reshape(12, 1).expand(12, cfg["B"]).contiguous()
        for _ in range(a.launches):
            r = eng.rollout(s0, d, t, wl.DT, 100, hold=1, cost_ref=cfg["cost_ref"])
            eng.argmin(r.cost)
    torch.cuda.synchronize()
    print("done", a.what)


if __name__ == "__main__":
    main()
