"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place ``/root/reference`` exists):

    python -m oracle.make_golden [--only planar|rollout|collision|closedloop|spiral_opt|lattice|tracking|plan_invalid] [--frames 400]

The reference has no tests or golden vectors of its own, so these files -- outputs of
the literal reference code on seeded inputs -- are what pins the oracle (``tests/test_oracle_pinned.py``)
and, through it, the CUDA path.  Every array is float64/bool/int exactly as the reference produced it.
Host facts are recorded in each file (numpy / scipy versions, ``np.linalg.norm`` closed form).
"""
from __future__ import annotations

import argparse
import itertools
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import collision_numpy, ref_loader  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DT = 1e-4


def _host_facts():
    import scipy
    return dict(numpy_version=np.__version__, scipy_version=scipy.__version__,
                norm2_mode=-1 if collision_numpy.probe_norm2_mode() is None else collision_numpy.probe_norm2_mode())


# --------------------------------------------------------------------------------------------- planar
def gen_planar():
    ref = ref_loader.load()
    VM = ref.vehicle_model
    vm = VM.VehicleModel(2.906, np.deg2rad(30), DT)
    p = VM.VehicleParameters()
    rng = np.random.default_rng(wl.SEED + 1)
    n = 96
    U = rng.uniform(3.0, 45.0, n)
    states = np.stack([U, rng.uniform(-2, 2, n), rng.uniform(-1, 1, n)]
                      + [U / p.rw * (1 + rng.uniform(-0.2, 0.2, n)) for _ in range(4)]
                      + [rng.uniform(-7, 7, n), rng.uniform(-100, 100, n), rng.uniform(-100, 100, n)], axis=1)
    tq = rng.uniform(-500, 500, (n, 4))
    mu = rng.uniform(0.3, 1.2, (n, 4))
    dl = np.concatenate([rng.uniform(-0.5, 0.5, (n, 2)), rng.uniform(-0.05, 0.05, (n, 2))], axis=1)
    axp, ayp = rng.uniform(-3, 3, n), rng.uniform(-6, 6, n)
    # fixed known-answer cases: KAT1 and the zero-slip state
    states[0] = [20.0, 0.5, 0.2, 66.0, 66.5, 65.0, 65.5, 0.3, 10.0, -5.0]
    tq[0], mu[0], dl[0], axp[0], ayp[0] = [50, 60, 70, 80], [1.0, 0.9, 0.8, 0.7], [0.05, 0.04, 0.01, -0.01], 0.4, -0.7
    states[1] = [25, 0, 0, 25 / p.rw, 25 / p.rw, 25 / p.rw, 25 / p.rw, 0, 0, 0]
    tq[1], mu[1], dl[1], axp[1], ayp[1] = [0] * 4, [1.0] * 4, [0] * 4, 0.0, 0.0
    sdot, misc, outs, Dafter = [], [], [], []
    for i in range(n):
        r = vm.planar_model(list(states[i]), list(tq[i]), list(mu[i]), list(dl[i]), p, axp[i], ayp[i])
        sdot.append(np.asarray(r[0], float))
        misc.append([r[1], r[2], r[3], r[4], r[6], r[7]])
        outs.append(np.asarray(r[5], float))
        Dafter.append([p.DFL, p.DFR, p.DRL, p.DRR])
    # single RK4 steps with the full return list
    rk_state, rk_sdot, rk_out, rk_axay = [], [], [], []
    for i in range(n):
        r = vm.planar_model_RK4(list(states[i]), list(tq[i]), list(mu[i]), list(dl[i]), p, axp[i], ayp[i])
        rk_state.append(np.asarray(r[0], float))
        rk_sdot.append(np.asarray(r[5], float))
        rk_out.append(np.asarray(r[6], float))
        rk_axay.append([r[7], r[8]])
    pr = VM.VehicleParameters()
    params = {k: float(getattr(pr, k)) for k in ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw",
                                                  "BFL", "CFL", "DFL")}
    np.savez_compressed(os.path.join(GOLDEN, "planar_model.npz"), states=states, torque=tq, mu_max=mu, delta=dl,
                        ax_prev=axp, ay_prev=ayp, state_dot=np.array(sdot), misc=np.array(misc),
                        outputs=np.array(outs), D_after=np.array(Dafter), rk4_state=np.array(rk_state),
                        rk4_state_dot=np.array(rk_sdot), rk4_outputs=np.array(rk_out), rk4_axay=np.array(rk_axay),
                        dt=DT, **{"param_" + k: v for k, v in params.items()}, **_host_facts())
    print("planar_model.npz:", n, "cases")


# -------------------------------------------------------------------------------------------- rollout
_CHECK_STEPS = (1, 10, 100, 250, 500)


def _rollout_worker(args):
    idx, state0, delta, torque, n_steps, hold = args
    ref = ref_loader.load()
    VM = ref.vehicle_model
    vm = VM.VehicleModel(2.906, np.deg2rad(30), DT)
    p = VM.VehicleParameters()
    out = np.zeros((len(idx), len(_CHECK_STEPS), 12))
    for k in range(len(idx)):
        st = list(state0[:10, k])
        ax, ay = state0[10, k], state0[11, k]
        for n in range(n_steps):
            d, t = delta[n // hold, 0, k], torque[n // hold, 0, k]
            r = vm.planar_model_RK4(st, [t, t, t, t], [1.0, 1.0, 1.0, 1.0], [d, d, 0, 0], p, ax, ay)
            st, ax, ay = r[0], r[7], r[8]
            if (n + 1) in _CHECK_STEPS:
                out[k, _CHECK_STEPS.index(n + 1), :10] = st
                out[k, _CHECK_STEPS.index(n + 1), 10:] = (ax, ay)
    return out


def gen_rollout(n_sub=256, n_steps=500):
    state0, delta, torque = wl.config2_rollouts(B=65536, n_steps=n_steps)
    # a fixed, spread-out subsample of the config-2 batch (same seed as the GPU run)
    idx = np.arange(n_sub) * (65536 // n_sub) + 7
    s0, d, t = state0[:, idx], delta[:, :, idx], torque[:, :, idx]
    nproc = min(mp.cpu_count(), 8)
    chunks = np.array_split(np.arange(n_sub), nproc)
    t0 = time.time()
    with mp.Pool(nproc) as pool:
        res = pool.map(_rollout_worker, [(c, s0[:, c], d[:, :, c], t[:, :, c], n_steps, wl.HOLD) for c in chunks])
    states = np.concatenate(res, axis=0)
    el = time.time() - t0
    np.savez_compressed(os.path.join(GOLDEN, "rollout_cfg2_sub.npz"), index=idx, state0=s0, delta=d, torque=t,
                        check_steps=np.array(_CHECK_STEPS), states=states, dt=DT, hold=wl.HOLD, seed=wl.SEED,
                        literal_steps_per_s=n_sub * n_steps / el, literal_procs=nproc, **_host_facts())
    print(f"rollout_cfg2_sub.npz: {n_sub} rollouts x {n_steps} steps, literal reference {n_sub*n_steps/el:.0f} steps/s "
          f"on {nproc} procs")


# ------------------------------------------------------------------------------------------ collision
def _collision_worker(args):
    paths, obstacles = args
    ref = ref_loader.load()
    cc = ref.collision_checker.CollisionChecker(list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT)
    return [bool(cc.collision_check(p, obstacles)) for p in paths]


def gen_collision(n_sub=48):
    ref = ref_loader.load()
    CC = ref.collision_checker.CollisionChecker
    cc = CC(list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT)
    w = wl.config3_lattice()
    idx = np.arange(n_sub) * (4096 // n_sub) + 3
    obstacles = w["obstacles"].tolist()
    paths = [[w["px"][i].tolist(), w["py"][i].tolist(), w["pyaw"][i].tolist()] for i in idx]
    nproc = min(mp.cpu_count(), 8)
    t0 = time.time()
    with mp.Pool(nproc) as pool:
        res = pool.map(_collision_worker, [(paths[k::nproc], obstacles) for k in range(nproc)])
    flags = np.zeros(n_sub, dtype=bool)
    for k in range(nproc):
        flags[k::nproc] = res[k]
    el = time.time() - t0
    print(f"collision literal: {n_sub} paths vs {len(obstacles)} points in {el:.1f}s, free={flags.mean():.2f}")

    # KAT3 (Appendix C) straight path cases
    kat_path = [[float(i) for i in range(1, 50)], [0.0] * 49, [0.0] * 50]
    kat_obs = [[[10, 1.6], [20, -1.6]], [[10, 1.6], [20, -1.4]], [[10, 1.5]], [[53.4, 0]], [[53.6, 0]], []]
    kat_flags = [bool(cc.collision_check(kat_path, o)) for o in kat_obs]

    # boundary stress: single obstacle points a few ulps either side of a circle's radius
    rng = np.random.default_rng(wl.SEED + 3)
    nb = 3000
    bx, by, byaw = rng.uniform(-100, 100, nb), rng.uniform(-100, 100, nb), rng.uniform(-np.pi, np.pi, nb)
    k = rng.integers(0, 3, nb)
    ang = rng.uniform(-np.pi, np.pi, nb)
    off = np.array(wl.CIRCLE_OFFSETS)[k]
    cx, cy = bx + off * np.cos(byaw), by + off * np.sin(byaw)
    ox, oy = cx + 1.5 * np.cos(ang), cy + 1.5 * np.sin(ang)
    for _ in range(2):   # jitter by a few ulps
        j = rng.integers(-3, 4, nb)
        ox = np.where(j > 0, np.nextafter(ox, np.inf), np.where(j < 0, np.nextafter(ox, -np.inf), ox))
    bflags = np.array([bool(cc.collision_check([[bx[i]], [by[i]], [byaw[i]]], [[ox[i], oy[i]]])) for i in range(nb)])
    print(f"boundary stress: {nb} cases, free={bflags.mean():.3f}")

    # select_best_path_index: literal on sub-lattices of config 3 (flags from the bit-exact oracle)
    free_all = collision_numpy.collision_check_batch(w["px"][:512], w["py"][:512], w["pyaw"][:512], w["obstacles"],
                                                     wl.CIRCLE_OFFSETS, wl.CIRCLE_RADII)
    sel_cases = []
    for lo, n in ((0, 7), (7, 7), (0, 64), (64, 200), (0, 512)):
        ps = [[w["px"][i].tolist(), w["py"][i].tolist(), w["pyaw"][i].tolist()] for i in range(lo, lo + n)]
        fl = [bool(f) for f in free_all[lo:lo + n]]
        best = cc.select_best_path_index(ps, fl, w["goal"])
        sel_cases.append((lo, n, -1 if best is None else best))
    # KAT4 symmetric lattice with exact ties
    k4_paths = [[[float(i) for i in range(1, 50)], [yv] * 49, [0.0] * 49] for yv in (-4.0, -2.0, 0.0, 2.0, 4.0)]
    k4_flags = [[True, True, False, True, True], [True] * 5, [False] * 5, [True, False, False, False, True]]
    k4_best = [cc.select_best_path_index(k4_paths, f, [49, 0, 25]) for f in k4_flags]
    k4_best = [-1 if b is None else b for b in k4_best]

    np.savez_compressed(os.path.join(GOLDEN, "collision_cfg3_sub.npz"), index=idx, px=w["px"][idx], py=w["py"][idx],
                        pyaw=w["pyaw"][idx], obstacles=w["obstacles"], free=flags,
                        offsets=np.array(wl.CIRCLE_OFFSETS), radii=np.array(wl.CIRCLE_RADII),
                        kat3_obstacles=np.array([np.array(o, float).reshape(-1, 2).tolist() + [[np.nan, np.nan]] * (2 - len(o))
                                                 for o in kat_obs]),
                        kat3_counts=np.array([len(o) for o in kat_obs]), kat3_free=np.array(kat_flags),
                        bnd_x=bx, bnd_y=by, bnd_yaw=byaw, bnd_ox=ox, bnd_oy=oy, bnd_free=bflags,
                        sel_cases=np.array(sel_cases), sel_free512=free_all, goal=np.array(w["goal"]),
                        weight=wl.PATH_SELECT_WEIGHT, kat4_flags=np.array(k4_flags), kat4_best=np.array(k4_best),
                        literal_paths_per_s=n_sub / el, literal_procs=nproc, seed=wl.SEED, **_host_facts())
    print("collision_cfg3_sub.npz written; select cases:", sel_cases, "kat4:", k4_best, "kat3:", kat_flags)


# ----------------------------------------------------------------------------------------- closed loop
class _SerialPool:
    """In-process stand-in for the reference's ``multiprocessing.Pool`` (local_planner.py:15, 369-374):
    same constructor/``starmap`` contract incl. ``ValueError`` for ``processes < 1``; results identical."""

    def __init__(self, processes=None):
        if processes is not None and processes < 1:
            raise ValueError("Number of processes must be at least 1")

    def starmap(self, fn, it):
        return [fn(*a) for a in it]


def gen_closedloop(frames=400, path_every=4):
    ref = ref_loader.load()
    drive, lp, cchk = ref.drive, ref.local_planner, ref.collision_checker
    lp.ThreadPool = _SerialPool
    rec = dict(delta=[], torque=[], state=[], axay=[], sdot=[], outputs=[])
    plan = dict(paths=[], frames=[], flags=[], best=[], goal=[], npaths=[])
    VM = drive.VehicleModel
    orig_rk4 = VM.planar_model_RK4
    orig_sel = cchk.CollisionChecker.select_best_path_index
    cur = {"frame": 0}

    def rk4(self, state, tire_torques, mu_max, delta, p, ax_prev, ay_prev):
        r = orig_rk4(self, state, tire_torques, mu_max, delta, p, ax_prev, ay_prev)
        rec["delta"].append(delta[0])
        rec["torque"].append(tire_torques[0])
        rec["state"].append(np.asarray(r[0], float))
        rec["axay"].append((r[7], r[8]))
        rec["sdot"].append(np.asarray(r[5], float))
        rec["outputs"].append(np.asarray(r[6], float))
        return r

    def sel(self, paths, flags, goal_state):
        b = orig_sel(self, paths, flags, goal_state)
        f = cur["frame"]
        plan["npaths"].append(len(paths))
        fl = np.ones(7, dtype=bool)
        fl[:len(flags)] = flags
        plan["flags"].append(fl)
        plan["best"].append(-1 if b is None else b)
        plan["goal"].append(goal_state[:2])
        if f % path_every == 0 and len(paths) == 7:
            plan["frames"].append(f)
            plan["paths"].append(np.array([[pp[0], pp[1], pp[2]] for pp in paths]))
        return b

    VM.planar_model_RK4 = rk4
    cchk.CollisionChecker.select_best_path_index = sel
    try:
        world = ref.env.world
        path = world.path
        car = drive.Car(path.px[10], path.py[10], path.pyaw[10], path.px, path.py, path.pyaw, DT)
        state0 = np.array(list(car.state) + [0.0, 0.0])
        frame_end = []
        t0 = time.time()
        for f in range(frames):
            cur["frame"] = f
            car.drive(f)
            frame_end.append([car.x, car.y, car.yaw, car.v, car.delta])
            if f % 20 == 0:
                print(f"frame {f}/{frames}  {time.time()-t0:.0f}s  x={car.x:.3f} y={car.y:.3f}", flush=True)
    finally:
        VM.planar_model_RK4 = orig_rk4
        cchk.CollisionChecker.select_best_path_index = orig_sel
    state = np.array(rec["state"])
    n = len(state)
    ctrl = slice(0, n, 10)
    # controls are constant over each block of 10 sub-steps (drive.py:128) -> store at control rate
    delta = np.array(rec["delta"], float)
    torque = np.array(rec["torque"], float)
    assert all(np.all(delta[i:i + 10] == delta[i]) and np.all(torque[i:i + 10] == torque[i]) for i in range(0, n, 10))
    keep = np.arange(9, n, 10)            # state after every 10th sub-step
    np.savez_compressed(os.path.join(GOLDEN, "closedloop_cfg1.npz"), state0=state0, delta=delta[ctrl], torque=torque[ctrl],
                        state_every10=state[keep], axay_every10=np.array(rec["axay"])[keep],
                        first_frame_states=state[:100], first_frame_sdot=np.array(rec["sdot"])[:100],
                        first_frame_outputs=np.array(rec["outputs"])[:100], frame_end=np.array(frame_end),
                        obstacle_xy=np.array(world.obstacle_xy), plan_flags=np.array(plan["flags"]),
                        plan_best=np.array(plan["best"]), plan_goal=np.array(plan["goal"], float),
                        plan_npaths=np.array(plan["npaths"]), plan_path_frames=np.array(plan["frames"]),
                        plan_paths=np.array(plan["paths"]), dt=DT, hold=10, frames=frames, **_host_facts())
    print(f"closedloop_cfg1.npz: {frames} frames, {n} RK4 steps, {time.time()-t0:.0f}s")



# ------------------------------------------------------------------------- closed-loop tracking (N3 / N4)
def gen_tracking(frames=200, every=8):
    """Per-frame pins for the batched Stanley/PID closed loop and the 45-column log:
    the unmodified reference Car (drive.py:112-154) runs `frames` frames; for every frame the waypoints the
    planner handed to the tracker (local_planner.py:419), the vehicle + controller state at the frame start and
    the 100 DataLog rows (drive.py:145-151) are recorded."""
    ref = ref_loader.load()
    drive, lp = ref.drive, ref.local_planner
    lp.ThreadPool = _SerialPool
    world = ref.env.world
    path = world.path
    car = drive.Car(path.px[10], path.py[10], path.pyaw[10], path.px, path.py, path.pyaw, DT)
    wps, starts, ctrl0, target_ids = [], [], [], []
    tracker = car.lateral_tracker
    orig_stanley = tracker.stanley_control
    ids = []

    def stanley(x, y, yaw, v):
        r = orig_stanley(x, y, yaw, v)
        ids.append(r[1])
        return r

    tracker.stanley_control = stanley
    t0 = time.time()
    kept = []
    for f in range(frames):
        keep = f % every == 0
        if keep:
            kept.append(f)
            starts.append(np.array(list(car.state) + [car.ax_prev, car.ay_prev], float))
            ctrl0.append([car.x_del[-1], car.total_vel_error, car.prev_vel, car.delta, car.torque_vec[0]])
        n_ids = len(ids)
        car.drive(f)
        if keep:
            wp = np.asarray(tracker._waypoints, float)  # set at sub-step 0 of this frame, used for all 10 updates
            wps.append(wp[:, :2].copy())
            target_ids.append(ids[n_ids:])
            print(f"tracking frame {f}/{frames} {time.time()-t0:.0f}s  waypoints {len(wp)}", flush=True)
    W = max(len(w) for w in wps)
    wp_pad = np.full((len(kept), W, 2), np.nan)
    for f, w in enumerate(wps):
        wp_pad[f, :len(w)] = w
    log = car.DataLog[:frames * 100].reshape(frames, 100, 45)[kept].copy()
    np.savez_compressed(os.path.join(GOLDEN, "tracking_frames.npz"), waypoints=wp_pad,
                        n_waypoints=np.array([len(w) for w in wps]), start=np.array(starts), ctrl0=np.array(ctrl0, float),
                        log=log, target_ids=np.array(target_ids), frame_index=np.array(kept), target_vel=car.target_vel,
                        gains=np.array([car.k, car.ksoft, car.max_steer, car.k_v, car.k_i, car.k_d], float),
                        lookahead=tracker._lookahead_distance, deadband=tracker.cross_track_deadband,
                        steer_filter=1e-5 / (2 * 0.001), dt=DT, frames=len(kept), **_host_facts())
    print(f"tracking_frames.npz: {len(kept)} of {frames} frames, W<= {W}, {time.time()-t0:.0f}s")


# ------------------------------------------------------------------ spiral sampling + frame transform (N1)
def gen_lattice(n=96):
    """Literal ``PathOptimizer.sample_spiral`` (path_optimizer.py:131-174) and ``transform_paths``
    (local_planner.py:424-470) on random optimisation parameters / ego poses."""
    ref = ref_loader.load()
    po = ref.path_optimizer.PathOptimizer()
    rng = np.random.default_rng(wl.SEED + 11)
    k1, k2 = rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n)
    sf = rng.uniform(20.0, 40.0, n)
    ego = np.stack([rng.uniform(0, 100, n), rng.uniform(0, 100, n), rng.uniform(-np.pi, np.pi, n)], 1)
    k1[:4], k2[:4] = 0.0, [0.0, 0.01, -0.02, 0.0]        # straight and one-sided cases
    x, y, t, gx, gy, gt = [], [], [], [], [], []
    for i in range(n):
        sp = po.sample_spiral([k1[i], k2[i], sf[i]])
        assert (len(sp[0]), len(sp[1]), len(sp[2])) == (49, 49, 50)
        tp = ref.local_planner.transform_paths([sp], list(ego[i]) + [25.0])[0]
        assert (len(tp[0]), len(tp[1]), len(tp[2])) == (49, 49, 49)
        x.append(sp[0]); y.append(sp[1]); t.append(sp[2])
        gx.append(tp[0]); gy.append(tp[1]); gt.append(tp[2])
    np.savez_compressed(os.path.join(GOLDEN, "lattice_paths.npz"), kappa1=k1, kappa2=k2, sf=sf, ego=ego,
                        x=np.array(x), y=np.array(y), t=np.array(t), gx=np.array(gx), gy=np.array(gy), gt=np.array(gt),
                        **_host_facts())
    print(f"lattice_paths.npz: {n} spirals")


# --------------------------------------------------------------------------- spiral optimisation (N2)
def gen_spiral_opt(n_eval=256, n_goals=192):
    """Literal ``PathOptimizer.objective`` / ``objective_grad`` (path_optimizer.py:183-530) at random points, and the
    literal ``optimize_spiral`` (scipy L-BFGS-B, :31-88) on planner-like and wider goal states: the optimiser's final
    parameters are captured from the ``sample_spiral`` call it ends with."""
    ref = ref_loader.load()
    po = ref.path_optimizer.PathOptimizer()
    rng = np.random.default_rng(wl.SEED + 13)
    ev_p = np.stack([rng.uniform(-0.2, 0.2, n_eval), rng.uniform(-0.2, 0.2, n_eval), rng.uniform(5.0, 60.0, n_eval)], 1)
    ev_goal = np.stack([rng.uniform(5.0, 50.0, n_eval), rng.uniform(-12.0, 12.0, n_eval), rng.uniform(-0.8, 0.8, n_eval)], 1)
    ev_f, ev_g = np.empty(n_eval), np.empty((n_eval, 3))
    for i in range(n_eval):
        po._xf, po._yf, po._tf = ev_goal[i]
        ev_f[i] = po.objective(list(ev_p[i]))
        ev_g[i] = po.objective_grad(list(ev_p[i]))
    # goal states: the planner's own pattern (30 m look-ahead, 7 lateral offsets of 2 m, small heading, local_planner.py
    # :154-275) plus a wider random set
    goals = []
    for k in range(n_goals // 2):
        gt = rng.uniform(-0.35, 0.35)
        gx, gy = rng.uniform(20.0, 40.0), rng.uniform(-3.0, 3.0)
        off = (k % 7 - 3) * 2.0
        goals.append([gx + off * np.cos(gt + np.pi / 2), gy + off * np.sin(gt + np.pi / 2), gt])
    for k in range(n_goals - len(goals)):
        goals.append([rng.uniform(8.0, 60.0), rng.uniform(-15.0, 15.0), rng.uniform(-1.0, 1.0)])
    goals = np.array(goals)
    captured = {}
    orig = po.sample_spiral

    def spy(p):
        captured["p"] = np.array(p, float)
        return orig(p)

    po.sample_spiral = spy
    res_p, res_f, res_end = np.empty((n_goals, 3)), np.empty(n_goals), np.empty((n_goals, 3))
    t0 = time.time()
    for i, g in enumerate(goals):
        sp = po.optimize_spiral(g[0], g[1], g[2])
        res_p[i] = captured["p"]
        res_f[i] = po.objective(list(captured["p"]))
        res_end[i] = [sp[0][-1], sp[1][-1], sp[2][-1]]
    el = time.time() - t0
    valid = np.array([np.linalg.norm(res_end[i] - goals[i]) <= 0.1 for i in range(n_goals)])     # local_planner.py:317-323
    # get_goal_state_set (local_planner.py:154-275) on random ego states / goal waypoints of a curved waypoint list
    lp = ref.local_planner.LocalPlanner(30, 7, 2, [-1.0, 1.0, 3.0], [1.5] * 3, 10, 1.0, 1.5, 2.0, 3.5)
    th = np.linspace(0.0, 1.2, 400)
    wps = np.stack([60.0 * np.sin(th), 60.0 * (1 - np.cos(th)), np.full(400, 25.0)], 1)
    gs_in, gs_out = [], []
    for k in range(48):
        gi = int(rng.integers(1, 400)) if k else 399                     # includes the last index (backward difference)
        ego = [wps[max(gi - 60, 0), 0] + rng.uniform(-2, 2), wps[max(gi - 60, 0), 1] + rng.uniform(-2, 2),
               rng.uniform(-np.pi, np.pi), 25.0]
        out = lp.get_goal_state_set(gi, list(wps[gi]), wps.tolist(), ego)
        gs_in.append([gi] + ego[:3])
        gs_out.append(out)
    np.savez_compressed(os.path.join(GOLDEN, "spiral_opt.npz"), goalset_waypoints=wps, goalset_in=np.array(gs_in),
                        goalset_out=np.array(gs_out), eval_p=ev_p, eval_goal=ev_goal, eval_f=ev_f, eval_grad=ev_g,
                        goals=goals, res_p=res_p, res_f=res_f, res_end=res_end, valid=valid,
                        literal_spirals_per_s=n_goals / el, **_host_facts())
    print(f"spiral_opt.npz: {n_eval} objective/gradient evaluations, {n_goals} optimisations in {el:.1f}s, valid {valid.mean():.2f}")


# ------------------------------------------------- planner core with dropped paths (plan_lattice)
def gen_plan_invalid():
    """Literal ``plan_paths`` (local_planner.py:277-325: a spiral whose end point misses its goal by more than 0.1 is
    DROPPED from the list) -> ``transform_paths`` (:424-470) -> literal ``collision_check`` per path ->
    ``select_best_path_index`` (collision_checker.py:134-203) on goal sets that contain unreachable goals at the head,
    in the middle, at the tail and at the would-be winner.  The chosen index refers to the FILTERED list."""
    ref = ref_loader.load()
    lp = ref.local_planner.LocalPlanner(30, 7, 2, list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT,
                                        1.0, 1.5, 2.0, 3.5)
    cc = lp._collision_checker
    rng = np.random.default_rng(wl.SEED + 21)
    unreachable = [(5.0, 3.0, 2.0), (4.0, 4.0, 0.0), (5.0, 3.0, 2.1), (4.0, 4.2, 0.0)]      # end-point error ~0.26 >> 0.1
    cases = []
    for c in range(8):
        n_goals = [7, 7, 9, 12, 7, 16, 7, 10][c]
        gt = rng.uniform(-0.3, 0.3)
        gx, gy = rng.uniform(24.0, 36.0), rng.uniform(-2.0, 2.0)
        goals = [[gx + (k - n_goals // 2) * 2.0 * np.cos(gt + np.pi / 2), gy + (k - n_goals // 2) * 2.0 * np.sin(gt + np.pi / 2),
                  gt, 25.0] for k in range(n_goals)]
        bad_at = [[0], [6], [3], [0, 5, 11], [2, 3, 4], [1, 8, 15], list(range(7)), []][c]
        for i, k in enumerate(bad_at):
            u = unreachable[(c + i) % len(unreachable)]
            goals[k] = [u[0], u[1], u[2], 25.0]
        ego = [rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-np.pi, np.pi), 25.0]
        paths, validity = lp.plan_paths(goals)
        paths = ref.local_planner.transform_paths(paths, ego)
        # obstacles: a box outline on the centre path's far half and a few scattered points, in the global frame
        obstacles = []
        if paths:
            mid = paths[(len(paths) // 2 + c) % len(paths)]
            cxo, cyo = mid[0][35], mid[1][35]
            X, Y = wl.box_outline(np.array([cxo + rng.uniform(-0.5, 0.5), cyo + rng.uniform(-0.5, 0.5)]), 1.0, 1.0, 0.21)
            obstacles = np.stack([X, Y], 1).tolist()
        flags = [bool(cc.collision_check(pth, obstacles)) for pth in paths]
        goal_state = [ego[0] + gx * np.cos(ego[2]) - gy * np.sin(ego[2]), ego[1] + gx * np.sin(ego[2]) + gy * np.cos(ego[2]), 25.0]
        best = cc.select_best_path_index(paths, flags, goal_state)
        cases.append(dict(goals=np.array(goals), ego=np.array(ego), obstacles=np.array(obstacles, float).reshape(-1, 2),
                          validity=np.array(validity, bool), flags=np.array(flags, bool), goal_state=np.array(goal_state),
                          best=-1 if best is None else int(best),
                          ends=np.array([[pth[0][-1], pth[1][-1]] for pth in paths], float).reshape(-1, 2)))
        print(f"case {c}: {n_goals} goals, valid {np.array(validity, int)}, free {np.array(flags, int)}, best(filtered) {best}")
    out = {"n_cases": len(cases), "weight": wl.PATH_SELECT_WEIGHT}
    for c, d in enumerate(cases):
        for k, v in d.items():
            out[f"c{c}_{k}"] = v
    np.savez_compressed(os.path.join(GOLDEN, "plan_invalid.npz"), **out, **_host_facts())
    print("plan_invalid.npz written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--frames", type=int, default=400)
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    if not ref_loader.available():
        raise SystemExit("reference not available here")
    if a.only in (None, "planar"):
        gen_planar()
    if a.only in (None, "rollout"):
        gen_rollout()
    if a.only in (None, "collision"):
        gen_collision()
    if a.only in (None, "closedloop"):
        gen_closedloop(a.frames)
    if a.only in (None, "spiral_opt"):
        gen_spiral_opt()
    if a.only in (None, "lattice"):
        gen_lattice()
    if a.only in (None, "tracking"):
        gen_tracking(min(a.frames, 200))
    if a.only in (None, "plan_invalid"):
        gen_plan_invalid()


if __name__ == "__main__":
    main()
