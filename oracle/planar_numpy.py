"""Vectorised NumPy oracle for the 7-DoF planar vehicle model.  TEST INFRASTRUCTURE ONLY.

Restates, for a batch of B independent vehicles held as arrays ``[.., B]``:
  * ``VehicleParameters``            reference ``libs/vehicle_model/vehicle_model.py:17-61``
  * ``VehicleModel.planar_model``    reference ``libs/vehicle_model/vehicle_model.py:220-425``
  * ``VehicleModel.planar_model_RK4`` reference ``libs/vehicle_model/vehicle_model.py:427-445``

The reference is scalar-only (an ndarray state raises at ``vehicle_model.py:309``); its
"NumPy path" is a Python loop over vehicles.  This file evaluates the same expressions in the same
operator order with float64 arrays, the ``if s != 0`` branch (``:309-348``) becoming ``np.where``.
Agreement with the literal reference is at the 1-ulp level but not bitwise: the scalar reference
computes ``sx ** 2`` through libm ``pow`` while arrays use ``x*x``.

Pinned by ``tests/golden/planar_*.npz`` (literal reference outputs, see ``make_golden.py``) and the
Appendix-C known-answer vectors; see ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

G = 9.81  # vehicle_model.py:230

# wheel order everywhere: FL, FR, RL, RR   (vehicle_model.py:224-227)
STATE_NAMES = ("U", "V", "wz", "wFL", "wFR", "wRL", "wRR", "yaw", "x", "y")


@dataclass
class VehicleParams:
    """Parameter block of reference ``VehicleParameters`` (vehicle_model.py:17-61).

    Constructor arguments and derived attributes carry the reference's names.  ``B*``/``C*``
    may be scalars or ``[B]`` arrays (config 5: one tyre set per rollout).
    """

    mf: float = 987.89
    mr: float = 869.93
    mus: float = 50
    L: float = 2.906
    ab_ratio: float = 0.85
    T: float = 1.536
    hg: float = 0.55419
    Jw: float = 1
    kf: float = 26290
    kr: float = 25830
    Efront: float = 0.0376
    Erear: float = 0
    LeverArm: float = 0.13256
    BFL: object = 20.6357
    CFL: object = 1.5047
    DFL: object = 1.1233
    rr: float = field(default=0.329, init=False)

    def __post_init__(self):
        self.m = self.mf + self.mr
        self.b = self.L / (1 + self.ab_ratio)
        self.a = self.L - self.b
        self.Izz = 0.5 * self.m * self.a * self.b
        self.rw = self.rr - (self.mf / 2 + self.mus) / self.kf
        self.wL = self.T / 2
        self.wR = self.T / 2
        for w in ("FR", "RL", "RR"):
            setattr(self, "B" + w, self.BFL)
            setattr(self, "C" + w, self.CFL)
            setattr(self, "D" + w, self.DFL)

    def Bvec(self):
        return [self.BFL, self.BFR, self.BRL, self.BRR]

    def Cvec(self):
        return [self.CFL, self.CFR, self.CRL, self.CRR]


def _rows(a, n, B):
    """Coerce a length-n list / [n] / [n,B] array to n rows broadcastable against [B]."""
    a = [np.asarray(r, dtype=np.float64) for r in a]
    assert len(a) == n
    return a


def planar_model(state, tire_torques, mu_max, delta, p: VehicleParams, ax_prev, ay_prev):
    """One right-hand-side evaluation for a batch.  Follows vehicle_model.py:220-425.

    state: 10 rows (each scalar or [B]); tire_torques, mu_max, delta: 4 rows; ax_prev, ay_prev: [B].
    Returns ``(state_dot[10,B], vx, vy, ax, ay, outputs[18,B], axc, ayc)`` like the reference list.
    """
    U, V, wz, wFL, wFR, wRL, wRR, yaw, x, y = _rows(state, 10, None)
    w = (wFL, wFR, wRL, wRR)
    dl = _rows(delta, 4, None)
    tq = _rows(tire_torques, 4, None)
    D = _rows(mu_max, 4, None)            # :232-235  mu_max replaces Pacejka D
    Bp = _rows(p.Bvec(), 4, None)
    Cp = _rows(p.Cvec(), 4, None)
    ax_prev = np.asarray(ax_prev, dtype=np.float64)
    ay_prev = np.asarray(ay_prev, dtype=np.float64)
    g = G

    # :245-253 static loads and load-transfer gains (parameter-only)
    fFz0 = p.b / (p.a + p.b) * p.m * g / 2
    fRz0 = p.a / (p.a + p.b) * p.m * g / 2
    DfzxL = p.m * p.hg * p.wR / ((p.a + p.b) * (p.wL + p.wR))
    DfzxR = p.m * p.hg * p.wL / ((p.a + p.b) * (p.wL + p.wR))
    DfzyF = p.m * p.hg * p.b / ((p.a + p.b) * (p.wL + p.wR))
    DfzyR = p.m * p.hg * p.a / ((p.a + p.b) * (p.wL + p.wR))
    # :255-258
    Fz = (fFz0 - DfzxL * ax_prev - DfzyF * ay_prev,
          fFz0 - DfzxR * ax_prev + DfzyF * ay_prev,
          fRz0 + DfzxL * ax_prev - DfzyR * ay_prev,
          fRz0 + DfzxR * ax_prev + DfzyR * ay_prev)

    # :261-271 wheel-centre velocities
    vxc = (U - p.T * wz / 2, U + p.T * wz / 2, U - p.T * wz / 2, U + p.T * wz / 2)
    vyc = (V + p.a * wz, V + p.a * wz, V - p.b * wz, V - p.b * wz)

    fx, fy, fxt, fyt, s_all = [], [], [], [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(4):
            cd, sd = np.cos(dl[i]), np.sin(dl[i])
            vx_t = vxc[i] * cd + vyc[i] * sd          # :274-281
            vy_t = -vxc[i] * sd + vyc[i] * cd
            sx = p.rw * w[i] / vx_t - 1               # :284-287
            sy = -vy_t / np.abs(vx_t)                 # :290-293
            s = np.sqrt(sx ** 2 + sy ** 2)            # :296-299
            mu = D[i] * np.sin(Cp[i] * np.arctan(Bp[i] * s))   # :303-306
            nz = s != 0                               # :309-348 (else-branch evaluates to 0)
            mux = np.where(nz, sx * mu / s, D[i] * np.sin(Cp[i] * np.arctan(Bp[i] * sx)))
            muy = np.where(nz, sy * mu / s, D[i] * np.sin(Cp[i] * np.arctan(Bp[i] * sy)))
            fxt_i = mux * Fz[i]                       # :351-360
            fyt_i = muy * Fz[i]
            fx.append(fxt_i * cd - fyt_i * sd)        # :363-373
            fy.append(fxt_i * sd + fyt_i * cd)
            fxt.append(fxt_i)
            fyt.append(fyt_i)
            s_all.append(s)

    # :376-385
    U_dot = 1 / p.m * (fx[0] + fx[1] + fx[2] + fx[3]) + V * wz
    V_dot = 1 / p.m * (fy[0] + fy[1] + fy[2] + fy[3]) - U * wz
    wz_dot = 1 / p.Izz * (p.a * (fy[0] + fy[1]) - p.b * (fy[2] + fy[3])
                          + p.T / 2 * (fx[1] - fx[0] + fx[3] - fx[2]))
    wFL_dot = (tq[0] - p.rw * fxt[0]) / p.Jw
    wFR_dot = (tq[1] - p.rw * fxt[1]) / p.Jw
    wRL_dot = (tq[2] - p.rw * fx[2]) / p.Jw          # chassis-frame force on the rear axle (:381-382)
    wRR_dot = (tq[3] - p.rw * fx[3]) / p.Jw
    cy, sy_ = np.cos(yaw), np.sin(yaw)
    yaw_dot = wz
    x_dot = U * cy - V * sy_
    y_dot = U * sy_ + V * cy

    shape = np.broadcast(U_dot, V_dot, wz_dot, wFL_dot, x_dot).shape
    bc = lambda r: np.broadcast_to(np.asarray(r, dtype=np.float64), shape)
    state_dot = np.stack([bc(r) for r in (U_dot, V_dot, wz_dot, wFL_dot, wFR_dot, wRL_dot, wRR_dot,
                                          yaw_dot, x_dot, y_dot)])
    # :410-416
    vx = U * cy - V * sy_
    vy = V * sy_ + U * cy                              # [sic] reference quirk 5
    axc = U_dot - V * wz
    ayc = V_dot + U * wz
    ax = axc * cy - ayc * sy_
    ay = axc * sy_ + ayc * cy
    outputs = np.stack([bc(r) for r in (*fx, *fy, *Fz, *s_all, fxt[0], fyt[0])])   # :420-423
    return state_dot, bc(vx), bc(vy), bc(ax), bc(ay), outputs, bc(axc), bc(ayc)


def planar_model_rk4(state, tire_torques, mu_max, delta, p, ax_prev, ay_prev, h):
    """One classic-RK4 step for a batch.  Follows vehicle_model.py:427-445 (same operator order).

    Returns ``(state_update[10,B], state_dot[10,B], outputs[18,B], axc[B], ayc[B])``.
    """
    state = np.asarray(state, dtype=np.float64)
    K1, _, _, _, _, o1, axc1, ayc1 = planar_model(state, tire_torques, mu_max, delta, p, ax_prev, ay_prev)
    K2, _, _, _, _, o2, axc2, ayc2 = planar_model(state + h / 2 * K1, tire_torques, mu_max, delta, p,
                                                  ax_prev, ay_prev)
    K3, _, _, _, _, o3, axc3, ayc3 = planar_model(state + h / 2 * K2, tire_torques, mu_max, delta, p,
                                                  ax_prev, ay_prev)
    K4, _, _, _, _, o4, axc4, ayc4 = planar_model(state + h * K3, tire_torques, mu_max, delta, p,
                                                  ax_prev, ay_prev)
    state_update = state + 1 / 6 * h * (K1 + 2 * K2 + 2 * K3 + K4)
    state_dot = (K1 + 2 * K2 + 2 * K3 + K4) / 6
    outputs = (o1 + 2 * o2 + 2 * o3 + o4) / 6
    axc = (axc1 + 2 * axc2 + 2 * axc3 + axc4) / 6
    ayc = (ayc1 + 2 * ayc2 + 2 * ayc3 + ayc4) / 6
    return state_update, state_dot, outputs, axc, ayc


def rollout(state0, delta_seg, torque_seg, p, dt, n_steps, hold=1, mu_max=None, ax0=None, ay0=None,
            store_stride=1, want_aux=False):
    """Open-loop rollout of a batch; the loop ``Car.drive`` runs per vehicle (drive.py:141-143).

    state0:     [10, B]
    delta_seg:  [n_seg, 4, B] or [n_seg, 1, B] (front steer on FL=FR, rear 0 as drive.py:143)
    torque_seg: [n_seg, 4, B] or [n_seg, 1, B] (equal torque on 4 wheels, stanley_controller.py:159)
    hold:       steps per control segment (zero-order hold; drive.py:128 uses 10)
    The returned ``axc, ayc`` of step n are step n+1's ``ax_prev, ay_prev`` (drive.py:141).
    Returns dict(traj[n_out,10,B], state_end[10,B], ax_end[B], ay_end[B], and with want_aux
    state_dot[n_out,10,B], outputs[n_out,18,B]).
    """
    state = np.array(state0, dtype=np.float64)
    Bn = state.shape[1]
    ax = np.zeros(Bn) if ax0 is None else np.array(ax0, dtype=np.float64)
    ay = np.zeros(Bn) if ay0 is None else np.array(ay0, dtype=np.float64)
    mu = [np.ones(Bn)] * 4 if mu_max is None else [np.broadcast_to(np.asarray(m, float), (Bn,)) for m in mu_max]
    zeros = np.zeros(Bn)
    traj, sdots, outs = [], [], []
    for n in range(n_steps):
        seg = n // hold
        d = delta_seg[seg]
        t = torque_seg[seg]
        delta = [d[0], d[0], zeros, zeros] if d.shape[0] == 1 else [d[0], d[1], d[2], d[3]]
        torque = [t[0]] * 4 if t.shape[0] == 1 else [t[0], t[1], t[2], t[3]]
        state, sd, out, ax, ay = planar_model_rk4(state, torque, mu, delta, p, ax, ay, dt)
        if store_stride and (n + 1) % store_stride == 0:
            traj.append(state.copy())
            if want_aux:
                sdots.append(sd)
                outs.append(out)
    res = dict(traj=np.stack(traj) if traj else np.zeros((0, 10, Bn)), state_end=state, ax_end=ax, ay_end=ay)
    if want_aux:
        res["state_dot"] = np.stack(sdots)
        res["outputs"] = np.stack(outs)
    return res
