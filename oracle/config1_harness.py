"""Runs config 1 -- ``animate.py``'s loop body (animate.py:27, 61-66 -> drive.py:112-154) -- on the UNMODIFIED
reference modules and records what ``tests/golden/closedloop_cfg1.npz`` holds.  Test infrastructure.

With ``use_engine=True`` the three seams are rebound by ``python_motionplanning_b200.install()`` first, so every
``planar_model_RK4`` call, the ``Pool.starmap`` collision fan-out and ``select_best_path_index`` run on the GPU engine
while ``Car.drive``, the planner, the controllers and the DataLog writes stay the reference's own code.
"""
from __future__ import annotations

import time

import numpy as np

DT = 1e-4          # animate.py:15-16: frame_dt / Veh_SIM_NUM = 0.01 / 100


def run(frames: int, use_engine: bool):
    from oracle import ref_loader
    ref = ref_loader.load()
    import python_motionplanning_b200 as mp
    rebound = []
    if use_engine:
        rebound = mp.install()
        assert len(rebound) == 3, rebound
        checker_cls = mp.CollisionChecker
    else:
        mp.uninstall()
        checker_cls = ref.collision_checker.CollisionChecker
    plan = dict(flags=[], best=[], npaths=[])
    orig_sel = checker_cls.select_best_path_index

    def sel(self, paths, flags, goal_state):
        b = orig_sel(self, paths, flags, goal_state)
        fl = np.ones(7, dtype=bool)
        fl[:len(flags)] = np.asarray(flags, dtype=bool)
        plan["flags"].append(fl)
        plan["best"].append(-1 if b is None else int(b))
        plan["npaths"].append(len(paths))
        return b

    checker_cls.select_best_path_index = sel
    try:
        world = ref.env.world
        path = world.path
        car = ref.drive.Car(path.px[10], path.py[10], path.pyaw[10], path.px, path.py, path.pyaw, DT)
        if use_engine:
            assert isinstance(car.kbm, mp.VehicleModel), "Car did not construct the GPU-backed VehicleModel"
            assert isinstance(car.local_motion_planner._collision_checker, mp.CollisionChecker)
        frame_end, frame_s = [], []
        for f in range(frames):
            t0 = time.perf_counter()
            paths, best_index, best_path = car.drive(f)
            frame_s.append(time.perf_counter() - t0)
            frame_end.append([car.x, car.y, car.yaw, car.v, car.delta])
    finally:
        checker_cls.select_best_path_index = orig_sel
        if use_engine:
            mp.uninstall()
    n = frames * 100
    return {"frame_end": np.array(frame_end), "datalog": car.DataLog[:n].copy(), "plan_flags": np.array(plan["flags"]),
            "plan_best": np.array(plan["best"]), "plan_npaths": np.array(plan["npaths"]), "frame_s": np.array(frame_s),
            "rebound": rebound, "kbm_class": type(car.kbm).__module__ + "." + type(car.kbm).__name__}


def compare(res, g, frames: int, tol: float = 1e-9):
    """Worst relative errors (|a - ref| / max(|ref|, 1)) against closedloop_cfg1.npz and the exact-match checks."""
    n = frames * 100
    rel = lambda a, r: float((np.abs(a - r) / np.maximum(np.abs(r), 1.0)).max())
    log = res["datalog"]
    out = {
        "frame_end": rel(res["frame_end"], g["frame_end"][:frames]),
        # DataLog columns 1:11 = state after every sub-step (drive.py:146); the golden keeps every 10th
        "state_every10": rel(log[9:n:10, 1:11], g["state_every10"][:n // 10]),
        "first_frame_states": rel(log[:100, 1:11], g["first_frame_states"]),
        "first_frame_sdot": rel(log[:100, 11:21], g["first_frame_sdot"]),
        "first_frame_outputs": rel(log[:100, 26:44], g["first_frame_outputs"]),
        # columns 21, 22:26 = filtered steering angle and the PID torques the reference's controllers produced from
        # the engine's states (drive.py:148-149); the golden stores them at control rate
        "delta": rel(log[0:n:10, 21], g["delta"][:n // 10]),
        "torque": rel(log[0:n:10, 22], g["torque"][:n // 10]),
        "time_column_exact": bool(np.array_equal(log[:, 0], np.arange(n) * DT)),
        "flags_exact": bool(np.array_equal(res["plan_flags"], g["plan_flags"][:frames])),
        "best_exact": bool(np.array_equal(res["plan_best"], g["plan_best"][:frames])),
        "npaths_exact": bool(np.array_equal(res["plan_npaths"], g["plan_npaths"][:frames])),
    }
    out["ok"] = all(v <= tol for k, v in out.items() if isinstance(v, float)) and all(
        v for v in out.values() if isinstance(v, bool))
    return out
