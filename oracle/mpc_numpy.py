"""NumPy restatement of the sampling-MPC control generator and cost.  TEST INFRASTRUCTURE ONLY.

MPC is a to-do in the reference (README.md:28), so there is no reference code for these two pieces;
this file restates the *documented definition* (include/b200mp.h) independently of the CUDA
code: Philox4x32-10 (Salmon et al., SC'11) keyed by the seed with counter (rollout, segment), one
Box-Muller pair per (rollout, segment); and the running cost
J = sum_n (x_n - xr_n)^2 + (y_n - yr_n)^2 + w_u (U_n - u_ref)^2 accumulated in step order.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        kk0 = np.uint64((k0 + r * _W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * _W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ kk0) & _MASK, lo1, (hi0 ^ c3 ^ kk1) & _MASK, lo0
    return c0, c1, c2, c3


def sample_controls(B, n_seg, seed, rollout0=0, delta_mean=0.0, delta_sigma=0.02, delta_clip=0.5235987755982988,
                    torque_mean=0.0, torque_sigma=50.0):
    """``delta[n_seg,1,B]``, ``torque[n_seg,1,B]`` as ``b200mp_mpc_sample_controls_f64`` defines them."""
    gid = np.arange(B, dtype=np.uint64) + np.uint64(rollout0)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    delta = np.empty((n_seg, 1, B))
    torque = np.empty((n_seg, 1, B))
    for seg in range(n_seg):
        c = philox4x32_10(gid & _MASK, gid >> np.uint64(32), np.full(B, seg, np.uint64), np.zeros(B, np.uint64), k0, k1)
        a = (c[0] << np.uint64(32)) | c[1]
        b = (c[2] << np.uint64(32)) | c[3]
        u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
        u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        rad = np.sqrt(-2.0 * np.log(u1))
        e0, e1 = rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)
        delta[seg, 0] = np.clip(delta_mean + delta_sigma * e0, -delta_clip, delta_clip)
        torque[seg, 0] = torque_mean + torque_sigma * e1
    return delta, torque


def rollout_cost(traj, cost_ref, w_u, u_ref):
    """traj [n,10,B] -> J[B], accumulated in step order with the kernel's grouping."""
    J = np.zeros(traj.shape[2])
    for n in range(traj.shape[0]):
        ex, ey, eu = traj[n, 8] - cost_ref[n, 0], traj[n, 9] - cost_ref[n, 1], traj[n, 0] - u_ref
        J = J + (ex * ex + ey * ey + w_u * (eu * eu))
    return J


def argmin_lowest(cost):
    """Lowest-index argmin with NaN treated as +inf; ``None`` when nothing is finite."""
    c = np.where(np.isnan(cost), np.inf, cost)
    if not np.isfinite(c).any():
        return None
    return int(np.argmin(c))
