"""ctypes front end of the plain-C oracle (``oracle/csrc/oracle.c``).  TEST INFRASTRUCTURE ONLY.

Array layouts are the same structure-of-arrays layouts the product C-ABI uses (rollout index
fastest), so a parity test hands identical host arrays to both sides.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


class OracleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw")] + [
        ("B", C.c_double * 4), ("C", C.c_double * 4), ("D", C.c_double * 4)]


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
        L.oracle_planar_model.argtypes = [dp, dp, dp, dp, C.POINTER(OracleParams), C.c_double, C.c_double, dp, dp, dp]
        L.oracle_planar_model.restype = None
        L.oracle_rollout.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, C.c_int, dp, C.c_int, C.c_int,
                                     dp, C.POINTER(OracleParams), ip, C.c_int, dp, dp, dp, dp, dp,
                                     C.c_double, C.c_double, C.c_int]
        L.oracle_rollout.restype = C.c_long
        L.oracle_collision_check.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, dp,
                                             C.c_int, up, dp, C.c_int]
        L.oracle_collision_check.restype = C.c_longlong
        L.oracle_select_best.argtypes = [C.c_int, dp, dp, up, C.c_double, C.c_double, C.c_double, C.c_int, dp, C.c_int]
        L.oracle_select_best.restype = C.c_int
        L.oracle_norm2.argtypes = [C.c_int, dp, dp, C.c_int, dp]
        L.oracle_norm2.restype = None
        L.oracle_max_threads.restype = C.c_int
        L.oracle_track_loop.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, dp, ip, C.c_int, C.c_int,
                                        C.POINTER(OracleParams), C.c_double, dp, C.c_int, C.c_int, dp, dp, ip, dp, dp,
                                        C.c_int]
        L.oracle_track_loop.restype = C.c_long
        L.oracle_stanley_control.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, C.c_int,
                                             ip, dp]
        L.oracle_stanley_control.restype = C.c_double
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_params(p, n_sets=None):
    """Pack a ``VehicleParams``-like object (reference attribute names) into an ``OracleParams`` array.

    ``BFL``.. / ``CFL``.. / ``DFL``.. may be scalars (one set) or ``[n_sets]`` arrays.
    """
    def col(name):
        return np.atleast_1d(np.asarray(getattr(p, name), dtype=np.float64))
    cols = {w: (col("B" + w), col("C" + w), col("D" + w)) for w in ("FL", "FR", "RL", "RR")}
    n = n_sets or max(max(len(c) for c in v) for v in cols.values())
    arr = (OracleParams * n)()
    for s in range(n):
        q = arr[s]
        for name in ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw"):
            setattr(q, name, float(getattr(p, name)))
        for i, w in enumerate(("FL", "FR", "RL", "RR")):
            b, c, d = cols[w]
            q.B[i] = float(b[s if len(b) > 1 else 0])
            q.C[i] = float(c[s if len(c) > 1 else 0])
            q.D[i] = float(d[s if len(d) > 1 else 0])
    return arr


def planar_model(state, tq, mu, delta, p, ax_prev, ay_prev):
    L = lib()
    par = make_params(p)
    sd, misc, out = np.zeros(10), np.zeros(6), np.zeros(18)
    L.oracle_planar_model(_dp(_f64(state)), _dp(_f64(tq)), _dp(_f64(mu)), _dp(_f64(delta)), par,
                          float(ax_prev), float(ay_prev), _dp(sd), _dp(misc), _dp(out))
    return sd, misc, out


def rollout(state0, delta, torque, params, dt, n_steps, hold=1, mu=None, param_set=None, store_stride=0,
            want_aux=False, ctrl_broadcast=False, cost_ref=None, w_u=0.1, u_ref=25.0, nthreads=None):
    """state0 [12,B]; delta [n_seg,dch,B]; torque [n_seg,tch,B] (``[n_seg,ch]`` when ctrl_broadcast)."""
    L = lib()
    state0 = _f64(state0)
    B = state0.shape[1]
    delta, torque = _f64(delta), _f64(torque)
    dch, tch = delta.shape[1], torque.shape[1]
    n_out = n_steps // store_stride if store_stride else 0
    traj = np.empty((n_out, 10, B)) if n_out else None
    aux = np.empty((n_out, 28, B)) if (n_out and want_aux) else None
    end = np.empty((12, B))
    cost = np.empty(B) if cost_ref is not None else None
    cref = _f64(cost_ref) if cost_ref is not None else None
    mu_a = _f64(mu) if mu is not None else None
    ps = np.ascontiguousarray(param_set, dtype=np.int32) if param_set is not None else None
    nthreads = nthreads or host_threads()
    L.oracle_rollout(B, n_steps, float(dt), int(hold), _dp(state0), _dp(delta), dch, _dp(torque), tch,
                     0 if ctrl_broadcast else 1, _dp(mu_a), params,
                     None if ps is None else ps.ctypes.data_as(C.POINTER(C.c_int)), int(store_stride),
                     _dp(traj), _dp(aux), _dp(end), _dp(cost), _dp(cref), float(w_u), float(u_ref), int(nthreads))
    return dict(traj=traj, aux=aux, state_end=end, cost=cost)


def collision_check(px, py, pyaw, obstacles, offsets, radii, early_exit=True, want_clearance=False, nthreads=None):
    L = lib()
    px, py = _f64(px), _f64(py)
    P, n = px.shape
    yaw = _f64(pyaw)[:, :n]
    pc, ps = _f64(np.cos(yaw)), _f64(np.sin(yaw))     # host numpy trig, as collision_checker.py:88-89
    obs = _f64(obstacles).reshape(-1, 2)
    off, rad = _f64(offsets), _f64(radii)
    free = np.empty(P, dtype=np.uint8)
    clr = np.empty(P) if want_clearance else None
    if want_clearance:
        early_exit = False
    tests = L.oracle_collision_check(P, n, len(off), _dp(off), _dp(rad), _dp(px), _dp(py), _dp(pc), _dp(ps),
                                     obs.shape[0], _dp(obs), 1 if early_exit else 0,
                                     free.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(clr),
                                     int(nthreads or host_threads()))
    return free.astype(bool), clr, int(tests)


def select_best(end_x, end_y, free, goal_xy, weight, norm_mode, nthreads=None):
    L = lib()
    ex, ey = _f64(end_x), _f64(end_y)
    fr = np.ascontiguousarray(free, dtype=np.uint8)
    scores = np.empty(len(ex))
    idx = L.oracle_select_best(len(ex), _dp(ex), _dp(ey), fr.ctypes.data_as(C.POINTER(C.c_ubyte)),
                               float(goal_xy[0]), float(goal_xy[1]), float(weight), int(norm_mode), _dp(scores),
                               int(nthreads or host_threads()))
    return (None if idx < 0 else int(idx)), scores


def norm2(v0, v1, mode):
    L = lib()
    v0, v1 = _f64(v0), _f64(v1)
    out = np.empty_like(v0)
    L.oracle_norm2(len(v0), _dp(v0), _dp(v1), int(mode), _dp(out))
    return out


def track_gains(k=100.0, k_soft=1.0, max_steer=np.deg2rad(30), kp=1000.0, ki=100.0, kd=0.0, lookahead=5.0,
                deadband=0.01, alpha=1e-5 / (2 * 0.001)):
    """[k, k_soft, max_steer, kp, ki, kd, lookahead, deadband, alpha] with the reference's values
    (drive.py:71-85, stanley_controller.py:44-45, drive.py:137)."""
    return np.array([k, k_soft, max_steer, kp, ki, kd, lookahead, deadband, alpha], dtype=np.float64)


def stanley_control(waypoints, x, y, yaw, v, gains, norm_mode):
    """One ``StanleyController.stanley_control`` call: (steer, target index, crosstrack error)."""
    L = lib()
    wp = _f64(waypoints).reshape(-1, 2)
    idx, cte = C.c_int(0), C.c_double(0.0)
    st = L.oracle_stanley_control(_dp(wp), len(wp), float(x), float(y), float(yaw), float(v), _dp(_f64(gains)),
                                  int(norm_mode), C.byref(idx), C.byref(cte))
    return st, idx.value, cte.value


def track_loop(state0, ctrl0, waypoints, wp_count, params, dt, n_steps, target_vel, gains, norm_mode, ctrl_every=10,
               vehicles_per_set=None, store_stride=0, want_log=False, step0=0, nthreads=None):
    """Closed-loop Stanley/PID tracking.  state0 [12,V]; ctrl0 [3,V]; waypoints [n_sets,Wmax,2]; wp_count [n_sets]."""
    L = lib()
    state0, ctrl0 = _f64(state0), _f64(ctrl0)
    V = state0.shape[1]
    wp = _f64(waypoints)
    if wp.ndim == 2:
        wp = wp[None]
    n_sets, Wmax = wp.shape[0], wp.shape[1]
    cnt = np.ascontiguousarray(wp_count if wp_count is not None else [Wmax] * n_sets, dtype=np.int32)
    vps = int(vehicles_per_set or -(-V // n_sets))
    n_out = n_steps // store_stride if store_stride else 0
    traj = np.empty((n_out, 10, V)) if n_out else None
    log = np.empty((n_out, 45, V)) if (n_out and want_log) else None
    n_ctrl = -(-n_steps // ctrl_every)
    tid = np.zeros((n_ctrl, V), dtype=np.int32)
    end, cend = np.empty((12, V)), np.empty((3, V))
    L.oracle_track_loop(V, int(n_steps), int(step0), float(dt), int(ctrl_every), _dp(state0), _dp(ctrl0), _dp(wp),
                        cnt.ctypes.data_as(C.POINTER(C.c_int)), Wmax, vps, params, float(target_vel), _dp(_f64(gains)),
                        int(norm_mode), int(store_stride), _dp(traj), _dp(log),
                        tid.ctypes.data_as(C.POINTER(C.c_int)), _dp(end), _dp(cend), int(nthreads or host_threads()))
    return dict(traj=traj, log=log, target_idx=tid, state_end=end, ctrl_end=cend)
