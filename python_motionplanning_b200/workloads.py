"""Seeded synthetic inputs for the BASELINE.json configurations.

Host-side NumPy only: these build the *inputs* that tests and ``bench.py`` hand to the CUDA path (and,
in tests, to the oracle).  Nothing here computes a result of the hot path.

Layouts are the structure-of-arrays device layouts of ``include/b200mp.h`` (rollout / path index
fastest where it matters).
"""
from __future__ import annotations

import numpy as np

SEED = 20261018
DT = 1e-4                      # animate.py:13-16 -> drive.py:109
HOLD = 10                      # controls are refreshed every 10 sub-steps, drive.py:128
CIRCLE_OFFSETS = (-1.0, 1.0, 3.0)   # drive.py:25
CIRCLE_RADII = (1.5, 1.5, 1.5)      # drive.py:26
PATH_SELECT_WEIGHT = 10.0           # drive.py:28
MAX_STEER = 0.5235987755982988      # np.deg2rad(30), drive.py:57
N_SPIRAL_SAMPLES = 50               # np.linspace default, path_optimizer.py:160

# default VehicleParameters() derived values (vehicle_model.py:17-61)
RW_DEFAULT = 0.329 - (987.89 / 2 + 50) / 26290


def config2_rollouts(B=65536, n_steps=500, hold=HOLD, seed=SEED, rw=RW_DEFAULT):
    """Config 2: open-loop batch.  Returns ``state0[12,B]``, ``delta[n_seg,1,B]``, ``torque[n_seg,1,B]``.

    U~U[5,40], V~U[-1,1], wz~U[-.5,.5], w_i=U/rw*(1+U[-.05,.05]), yaw~U[-pi,pi], x,y~U[-100,100],
    ax_prev=ay_prev=0; front steer ~U[-.1,.1] on FL=FR (rear 0), torque ~U[-300,300] equal on 4 wheels,
    piecewise constant over ``hold`` steps.
    """
    rng = np.random.default_rng(seed)
    U = rng.uniform(5.0, 40.0, B)
    rows = [U, rng.uniform(-1.0, 1.0, B), rng.uniform(-0.5, 0.5, B)]
    rows += [U / rw * (1.0 + rng.uniform(-0.05, 0.05, B)) for _ in range(4)]
    rows += [rng.uniform(-np.pi, np.pi, B), rng.uniform(-100.0, 100.0, B), rng.uniform(-100.0, 100.0, B)]
    rows += [np.zeros(B), np.zeros(B)]
    state0 = np.ascontiguousarray(np.stack(rows))
    n_seg = -(-n_steps // hold)
    delta = rng.uniform(-0.1, 0.1, (n_seg, 1, B))
    torque = rng.uniform(-300.0, 300.0, (n_seg, 1, B))
    return state0, delta, torque


def sample_spirals(kappa1, kappa2, sf):
    """Vectorised cubic-spiral sampling for ``[P]`` parameter arrays -> ``x[P,49], y[P,49], yaw[P,50]``.

    Same construction as the reference's ``sample_spiral`` (path_optimizer.py:131-174): 50 arc-length
    samples, heading polynomial, cumulative trapezoid without an initial value (49 points, 50 yaws).
    """
    p1 = np.asarray(kappa1, dtype=np.float64)
    p2 = np.asarray(kappa2, dtype=np.float64)
    sf = np.asarray(sf, dtype=np.float64)
    b = -(-9.0 * p1 + 9.0 * p2 / 2.0) / sf
    c = (-45.0 * p1 / 2.0 + 18.0 * p2) / sf ** 2
    d = -(-27.0 * p1 / 2.0 + 27.0 * p2 / 2.0) / sf ** 3
    s = np.linspace(0.0, sf, N_SPIRAL_SAMPLES, axis=-1)                 # [P,50]
    t = (b / 2)[:, None] * s ** 2 + (c / 3)[:, None] * s ** 3 + (d / 4)[:, None] * s ** 4
    ds = np.diff(s, axis=-1)
    ct, st = np.cos(t), np.sin(t)
    x = np.cumsum(ds * (ct[:, 1:] + ct[:, :-1]) / 2.0, axis=-1)
    y = np.cumsum(ds * (st[:, 1:] + st[:, :-1]) / 2.0, axis=-1)
    return x, y, t


def transform_to_global(x, y, yaw, ego_x, ego_y, ego_yaw):
    """Ego-frame paths -> global frame (local_planner.py:424-470); keeps the first ``x.shape[1]`` yaws."""
    n = x.shape[1]
    c, s = np.cos(ego_yaw)[:, None], np.sin(ego_yaw)[:, None]
    gx = ego_x[:, None] + x * c - y * s
    gy = ego_y[:, None] + x * s + y * c
    gyaw = yaw[:, :n] + ego_yaw[:, None]
    return gx, gy, gyaw


def box_outline(corner, width, length, ds):
    """Outline points of one box obstacle, four edges in the order of env.py:93-127."""
    x1, y1 = corner
    x2, y2 = x1 + length, y1 + width
    bottom = np.arange(x1, x2, ds)
    side = np.arange(y1, y2, ds)
    top = np.arange(x2, x1, -ds)
    X = np.concatenate([bottom, np.full(len(side), x2), top, np.full(len(side), x1)])
    Y = np.concatenate([np.full(len(bottom), y1), side, np.full(len(top), y2), side])
    return X, Y


def config3_lattice(P=4096, M=10000, seed=SEED, area=100.0, obstacle_margin=50.0):
    """Config 3: P spirals x 49 points (x, y, yaw each [P,49]) vs M obstacle points ``[M,2]``.

    Spirals: kappa1,kappa2~U[-.05,.05], sf~U[20,40]; ego poses x,y~U[0,area], yaw~U[-pi,pi].
    Obstacles: boxes (width 6, length 4.5, ds 0.21; env.py:181) at corners ~U[-margin, area+margin]^2
    (the margin keeps the collision-free fraction inside the 20-80 % band SURVEY.md §8d asks for; with
    margin 0 only 11 % of the paths are free), concatenated and truncated to exactly M points.  Returns dict with ``goal=[50,50,25]`` and ``weight=10``.
    """
    rng = np.random.default_rng(seed)
    k1 = rng.uniform(-0.05, 0.05, P)
    k2 = rng.uniform(-0.05, 0.05, P)
    sf = rng.uniform(20.0, 40.0, P)
    ex, ey = rng.uniform(0.0, area, P), rng.uniform(0.0, area, P)
    eyaw = rng.uniform(-np.pi, np.pi, P)
    x, y, t = sample_spirals(k1, k2, sf)
    gx, gy, gyaw = transform_to_global(x, y, t, ex, ey, eyaw)
    xs, ys, n = [], [], 0
    while n < M:
        X, Y = box_outline(rng.uniform(-obstacle_margin, area + obstacle_margin, 2), 6.0, 4.5, 0.21)
        xs.append(X)
        ys.append(Y)
        n += len(X)
    obs = np.stack([np.concatenate(xs)[:M], np.concatenate(ys)[:M]], axis=1) if M else np.zeros((0, 2))
    return dict(px=np.ascontiguousarray(gx), py=np.ascontiguousarray(gy), pyaw=np.ascontiguousarray(gyaw),
                obstacles=np.ascontiguousarray(obs), goal=[50.0, 50.0, 25.0], weight=PATH_SELECT_WEIGHT,
                offsets=list(CIRCLE_OFFSETS), radii=list(CIRCLE_RADII),
                spiral_params=np.stack([k1, k2, sf]), ego=np.stack([ex, ey, eyaw]))


def config4_mpc(B=1 << 20, n_steps=100, seed=SEED, rw=RW_DEFAULT, start=(0.5, 10.0, 0.0), v0=25.0):
    """Config 4: B control sequences from one shared start state (the config-1 start, drive.py:46,64-65).

    Control law: ``delta_t = clip(dbar + 0.02 eps_t, +-max_steer)``, ``tau_t = tbar + 50 eps'_t``,
    eps iid N(0,1) drawn on the device from a counter-based generator keyed (seed, rollout, step).
    Cost reference: straight line at v0 along the start heading.  Returns a dict of scalars/arrays.
    """
    x0, y0, yaw0 = start
    state0 = np.array([v0, 0.0, 0.0, v0 / rw, v0 / rw, v0 / rw, v0 / rw, yaw0, x0, y0, 0.0, 0.0])
    t = (np.arange(n_steps) + 1) * DT
    ref = np.stack([x0 + v0 * np.cos(yaw0) * t, y0 + v0 * np.sin(yaw0) * t], axis=1)
    return dict(B=B, n_steps=n_steps, state0=state0, delta_mean=0.0, delta_sigma=0.02, delta_clip=MAX_STEER,
                torque_mean=0.0, torque_sigma=50.0, seed=seed, cost_ref=np.ascontiguousarray(ref),
                w_u=0.1, u_ref=v0)


def config5_sweep(n_sets=256, n_man=4096, seed=SEED, rw=RW_DEFAULT):
    """Config 5: ``n_sets`` tyre-coefficient sets x ``n_man`` manoeuvres (set-major rollout order).

    Sets: B~U[8,25], C~U[1.2,1.9], D~U[.3,1.2] (same on 4 wheels, vehicle_model.py:41-54).
    Manoeuvres: U0~U[8,35], step steer ~U[-.08,.08], torque ~U[-200,200], constant over the rollout.
    Returns ``sets[n_sets,3]``, ``state0[12,B]``, ``delta[1,1,B]``, ``torque[1,1,B]``, ``param_set[B]``.
    """
    rng = np.random.default_rng(seed)
    sets = np.stack([rng.uniform(8.0, 25.0, n_sets), rng.uniform(1.2, 1.9, n_sets), rng.uniform(0.3, 1.2, n_sets)], 1)
    U0 = rng.uniform(8.0, 35.0, n_man)
    dl = rng.uniform(-0.08, 0.08, n_man)
    tq = rng.uniform(-200.0, 200.0, n_man)
    B = n_sets * n_man
    state0 = np.zeros((12, B))
    U = np.tile(U0, n_sets)
    state0[0] = U
    state0[3:7] = U / rw
    delta = np.tile(dl, n_sets)[None, None, :]
    torque = np.tile(tq, n_sets)[None, None, :]
    param_set = np.repeat(np.arange(n_sets, dtype=np.int32), n_man)
    return sets, state0, np.ascontiguousarray(delta), np.ascontiguousarray(torque), param_set


def tracking_fleet(V: int = 4096, n_sets: int = 4, W: int = 3000, ds: float = 0.01, seed: int = SEED + 7):
    """Closed-loop tracking workload: ``n_sets`` waypoint lists of ``W`` points at ``ds`` spacing
    (what the planner hands the tracker: its best path re-interpolated at 1 cm, local_planner.py:390-419) --
    clothoid-like arcs from random poses -- and ``V`` vehicles (``V // n_sets`` per list) starting near the head of
    their list with a lateral / heading / speed offset (speeds within 2 m/s of the 25 m/s target: the reference's
    PID gain of 1000 N m per m/s spins the wheels up for larger errors).  Returns ``state0 [12,V]``, ``waypoints [n_sets,W,2]``."""
    rng = np.random.default_rng(seed)
    wp = np.empty((n_sets, W, 2))
    vps = -(-V // n_sets)
    state0 = np.zeros((12, V))
    rw = 0.329 - (987.89 / 2 + 50) / 26290            # VehicleParameters().rw (vehicle_model.py:38)
    for s in range(n_sets):
        x0, y0, th0 = rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-np.pi, np.pi)
        k0, k1 = rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02)
        arc = np.arange(W) * ds
        th = th0 + k0 * arc + 0.5 * (k1 - k0) / (W * ds) * arc ** 2
        wp[s, :, 0] = x0 + np.concatenate([[0.0], np.cumsum(np.cos(th[:-1]) * ds)])
        wp[s, :, 1] = y0 + np.concatenate([[0.0], np.cumsum(np.sin(th[:-1]) * ds)])
        lo, hi = s * vps, min(V, (s + 1) * vps)
        n = hi - lo
        if n <= 0:
            continue
        along = rng.uniform(0.0, 2.0, n)               # up to 2 m into the list
        lat = rng.uniform(-0.5, 0.5, n)
        i0 = np.minimum((along / ds).astype(int), W - 2)
        yaw = th[i0] + rng.uniform(-0.1, 0.1, n)
        U = rng.uniform(23.0, 27.0, n)
        state0[0, lo:hi] = U
        state0[1, lo:hi] = rng.uniform(-0.3, 0.3, n)
        state0[2, lo:hi] = rng.uniform(-0.1, 0.1, n)
        state0[3:7, lo:hi] = (U / rw)[None, :] * (1 + rng.uniform(-0.02, 0.02, (4, n)))
        state0[7, lo:hi] = yaw
        state0[8, lo:hi] = wp[s, i0, 0] - lat * np.sin(th[i0])
        state0[9, lo:hi] = wp[s, i0, 1] + lat * np.cos(th[i0])
    return state0, wp
