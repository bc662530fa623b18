"""Drop-in ``VehicleParameters`` / ``VehicleModel`` with the reference's call surface, computed on the GPU.

Mirrors ``libs/vehicle_model/vehicle_model.py`` of the reference for the hot path only:
  * ``VehicleParameters(...)``                    (:17-61)  same constructor, same derived attributes
  * ``VehicleModel(wheelbase, max_steer, dt)``    (:69-95)
  * ``.planar_model(state, tire_torques, mu_max, delta, p, ax_prev, ay_prev)``      (:220-425)
  * ``.planar_model_RK4(same)``                   (:427-445)
Return lists have the reference's order and types (numpy arrays / floats); ``p.DFL..p.DRR`` equal ``mu_max``
after a call, as in the reference (:232-235).  The scalar calls run the same CUDA kernels with a batch of
one (one small H2D, one launch, one small D2H through pinned staging buffers), so ``Car.drive``
(drive.py:141-143) works unchanged once ``install()`` has rebound ``drive.VehicleModel``.

The dead legacy models of the same file (``kinematic_model``, ``bicycle_model``, ``planar_integrate``)
are out of scope and are not provided.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .engine import Engine, uploaded_token


class VehicleParameters:
    """Parameter block with the reference's constructor and attribute names (vehicle_model.py:17-61)."""

    def __init__(self, mf=987.89, mr=869.93, mus=50, L=2.906, ab_ratio=0.85, T=1.536, hg=0.55419, Jw=1,
                 kf=26290, kr=25830, Efront=0.0376, Erear=0, LeverArm=0.13256, BFL=20.6357, CFL=1.5047, DFL=1.1233):
        self.rr = 0.329
        self.mus, self.mf, self.mr = mus, mf, mr
        self.m = mf + mr
        self.L = L
        self.ab_ratio = ab_ratio
        self.b = self.L / (1 + self.ab_ratio)
        self.a = self.L - self.b
        self.Izz = 0.5 * self.m * self.a * self.b
        self.Jw, self.hg, self.T = Jw, hg, T
        self.kf, self.kr = kf, kr
        self.rw = self.rr - (self.mf / 2 + self.mus) / self.kf
        for w in ("FL", "FR", "RL", "RR"):
            setattr(self, "B" + w, BFL)
            setattr(self, "C" + w, CFL)
            setattr(self, "D" + w, DFL)
        self.Efront, self.Erear = Efront, Erear
        self.E = [Efront, Efront, Erear, Erear]
        self.LeverArm = LeverArm
        self.wL = self.T / 2
        self.wR = self.T / 2


_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    """The process-wide engine on the current CUDA device (raises without a GPU: no CPU fallback)."""
    global _default_engine
    if _default_engine is None:
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        _default_engine = Engine(dev)
    return _default_engine


class VehicleModel:
    """GPU-backed stand-in for the reference's ``VehicleModel`` (planar 7-DoF model only)."""

    def __init__(self, wheelbase=1.0, max_steer=0.7, dt=0.05, engine: Optional[Engine] = None):
        self.dt = dt
        self.wheelbase = wheelbase
        self.max_steer = max_steer
        self._engine = engine
        self._stage = None
        self._param_sig = None

    # -- staging: [state 12 | torque 4 | mu 4 | delta 4] in, [state_end 12 | aux 28] / [sd 10 | misc 6 | out 18] out
    def _buffers(self):
        if self._stage is None:
            eng = self._engine or default_engine()
            self._engine = eng
            self._h_in = torch.empty(24, dtype=torch.float64).pin_memory()
            self._h_out = torch.empty(40, dtype=torch.float64).pin_memory()
            self._d_in = eng.empty(24)
            self._d_out = eng.empty(40)
            self._np_in = self._h_in.numpy()
            self._np_out = self._h_out.numpy()
            self._stage = True
        return self._engine

    def _load_inputs(self, state, tire_torques, mu_max, delta, p, ax_prev, ay_prev):
        eng = self._buffers()
        # re-upload the parameter table only when a field the model reads has changed (D is replaced by
        # mu_max on every call and travels as the mu array instead)
        sig = (uploaded_token(eng.device), p.m, p.a, p.b, p.Izz, p.Jw, p.hg, p.T, p.wL, p.wR, p.rw,
               p.BFL, p.BFR, p.BRL, p.BRR, p.CFL, p.CFR, p.CRL, p.CRR)
        if sig != self._param_sig:
            eng.set_params(p)
            self._param_sig = (uploaded_token(eng.device),) + sig[1:]
        a = self._np_in
        a[0:10] = state
        a[10] = ax_prev
        a[11] = ay_prev
        a[12:16] = tire_torques
        a[16:20] = mu_max
        a[20:24] = delta
        # the reference overwrites the Pacejka peaks with mu_max on every call (:232-235)
        p.DFL, p.DFR, p.DRL, p.DRR = mu_max
        self._d_in.copy_(self._h_in, non_blocking=True)
        return eng

    def planar_model(self, state, tire_torques, mu_max, delta, p, ax_prev, ay_prev):
        """One right-hand-side evaluation -> ``[state_dot, vx, vy, ax, ay, outputs, axc, ayc]`` (:425)."""
        eng = self._load_inputs(state, tire_torques, mu_max, delta, p, ax_prev, ay_prev)
        d = self._d_in
        # the three results land directly in the staging buffer ([ax_prev, ay_prev] sit side by side in the input one)
        o = self._d_out
        eng.planar_model_batch(d[0:10].view(10, 1), d[12:16].view(4, 1), d[16:20].view(4, 1), d[20:24].view(4, 1), None, None,
                               axay=d[10:12].view(2, 1), out=(o[0:10].view(10, 1), o[10:16].view(6, 1), o[16:34].view(18, 1)))
        self._h_out.copy_(self._d_out, non_blocking=True)
        torch.cuda.current_stream(eng.tdev).synchronize()
        o = self._np_out
        f = np.float64
        return [o[0:10].copy(), f(o[10]), f(o[11]), f(o[12]), f(o[13]), o[16:34].copy(), f(o[14]), f(o[15])]

    def planar_model_RK4(self, state, tire_torques, mu_max, delta, p, ax_prev, ay_prev):
        """One RK4 step of size ``self.dt`` ->
        ``[state_update, x, y, yaw, U, state_dot, outputs, axc, ayc]`` (:445)."""
        eng = self._load_inputs(state, tire_torques, mu_max, delta, p, ax_prev, ay_prev)
        d = self._d_in
        # state_end and the 28 logged outputs land directly in the staging buffer: one launch, no device-side copies
        eng.rollout(d[0:12].view(12, 1), d[20:24].view(1, 4, 1), d[12:16].view(1, 4, 1), self.dt, 1, hold=1,
                    mu=d[16:20].view(4, 1), store_stride=1, want_aux=True, state_out=self._d_out[0:12].view(12, 1),
                    aux_out=self._d_out[12:40].view(1, 28, 1))
        self._h_out.copy_(self._d_out, non_blocking=True)
        torch.cuda.current_stream(eng.tdev).synchronize()
        o = self._np_out
        f = np.float64
        st = o[0:10].copy()
        return [st, st[8], st[9], st[7], st[0], o[12:22].copy(), o[22:40].copy(), f(o[10]), f(o[11])]
