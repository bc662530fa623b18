"""NumPy restatement of the cubic-spiral optimisation problem.  TEST INFRASTRUCTURE ONLY.

  * ``objective`` / ``objective_grad``   reference ``libs/motionplanner/path_optimizer.py:183-530``
    (``fbe + 25 (fxf + fyf) + 30 ftf`` and its gradient; the reference's functions are machine-generated symbolic
    expansions of the formulas below)
  * ``optimize``                          the problem ``optimize_spiral`` poses (:31-88): start ``[0, 0, |goal|]``,
    bounds ``p1, p2 in [-0.5, 0.5]``, ``sf >= |goal|``.  The reference hands it to scipy's L-BFGS-B, which is not
    part of the reference's sources; this restatement solves the same bounded problem with a projected
    Levenberg-Marquardt iteration to a tight tolerance, so parity with the reference is "same minimiser to the
    accuracy scipy stops at" (L-BFGS-B: ftol 2.2e-9, gtol 1e-5), never bitwise.

Formulas (p = [p0, p1, p2, p3, sf] with p0 = p3 = 0 fixed; u = s / sf in [0, 1]):
    theta(u) = sf * g(u),  g(u) = A u^2 / 2 - Bc u^3 / 3 - Cc u^4 / 4,
        A = 9 p1 - 4.5 p2,  Bc = 22.5 p1 - 18 p2,  Cc = -13.5 p1 + 13.5 p2
    x(sf) = sf / 24 * sum_i w_i cos theta(i / 8),  y(sf) likewise with sin,  w = [1 4 2 4 2 4 2 4 1]   (Simpson, 8 panels)
    fxf = (xf - x)^2,  fyf = (yf - y)^2,  ftf = (tf - theta(1))^2,  theta(1) = sf * 3 (p1 + p2) / 8
    fbe = sf * (324 p1^2 + 324 p2^2 - 81 p1 p2) / 840
"""
from __future__ import annotations

import numpy as np

W = np.array([1.0, 4.0, 2.0, 4.0, 2.0, 4.0, 2.0, 4.0, 1.0])
U = np.arange(9) / 8.0
G1 = 4.5 * U ** 2 - 7.5 * U ** 3 + 3.375 * U ** 4          # d g / d p1
G2 = -2.25 * U ** 2 + 6.0 * U ** 3 - 3.375 * U ** 4        # d g / d p2


def _parts(p, goal):
    p1, p2, sf = (np.asarray(v, dtype=np.float64)[..., None] for v in (p[..., 0], p[..., 1], p[..., 2]))
    g = p1 * G1 + p2 * G2                                   # [..., 9]
    th = sf * g
    c, s = np.cos(th), np.sin(th)
    X = sf[..., 0] / 24.0 * (W * c).sum(-1)
    Y = sf[..., 0] / 24.0 * (W * s).sum(-1)
    T = sf[..., 0] * g[..., -1]
    return p1[..., 0], p2[..., 0], sf[..., 0], g, c, s, goal[..., 0] - X, goal[..., 1] - Y, goal[..., 2] - T


def objective(p, goal):
    """``PathOptimizer.objective`` for ``p [..., 3]`` = (p1, p2, sf) and ``goal [..., 3]`` = (xf, yf, tf)."""
    p, goal = np.asarray(p, dtype=np.float64), np.asarray(goal, dtype=np.float64)
    p1, p2, sf, g, c, s, ex, ey, et = _parts(p, goal)
    fbe = sf * (324.0 * p1 * p1 + 324.0 * p2 * p2 - 81.0 * p1 * p2) / 840.0
    return fbe + 25.0 * (ex * ex + ey * ey) + 30.0 * et * et


def residual_jacobian(p, goal):
    """Residuals ``(ex, ey, et)`` and their Jacobian wrt (p1, p2, sf): ``[..., 3, 3]``."""
    p, goal = np.asarray(p, dtype=np.float64), np.asarray(goal, dtype=np.float64)
    p1, p2, sf, g, c, s, ex, ey, et = _parts(p, goal)
    sf_ = sf[..., None]
    J = np.empty(p.shape[:-1] + (3, 3))
    # d ex / d(.) = - dX / d(.)
    J[..., 0, 0] = sf * sf / 24.0 * (W * s * G1).sum(-1)
    J[..., 0, 1] = sf * sf / 24.0 * (W * s * G2).sum(-1)
    J[..., 0, 2] = -((W * c).sum(-1) / 24.0 - sf / 24.0 * (W * s * g).sum(-1))
    J[..., 1, 0] = -sf * sf / 24.0 * (W * c * G1).sum(-1)
    J[..., 1, 1] = -sf * sf / 24.0 * (W * c * G2).sum(-1)
    J[..., 1, 2] = -((W * s).sum(-1) / 24.0 + sf / 24.0 * (W * c * g).sum(-1))
    J[..., 2, 0] = -sf * G1[-1]
    J[..., 2, 1] = -sf * G2[-1]
    J[..., 2, 2] = -g[..., -1]
    del sf_
    return np.stack([ex, ey, et], -1), J


def objective_grad(p, goal):
    """``PathOptimizer.objective_grad``: gradient wrt (p1, p2, sf)."""
    p = np.asarray(p, dtype=np.float64)
    r, J = residual_jacobian(p, goal)
    wts = np.array([25.0, 25.0, 30.0])
    grad = 2.0 * np.einsum("...i,...ij->...j", r * wts, J)
    p1, p2, sf = p[..., 0], p[..., 1], p[..., 2]
    grad[..., 0] += sf * (648.0 * p1 - 81.0 * p2) / 840.0
    grad[..., 1] += sf * (648.0 * p2 - 81.0 * p1) / 840.0
    grad[..., 2] += (324.0 * p1 * p1 + 324.0 * p2 * p2 - 81.0 * p1 * p2) / 840.0
    return grad


def optimize(goal, max_iter=100, tol=1e-13):
    """Minimise the objective for one goal state from the reference's start point inside its bounds.

    Projected Levenberg-Marquardt: model Hessian ``2 J^T diag(w) J + Hessian(fbe)``, active bounds frozen when the
    gradient points outward, step accepted when the objective decreases.  Returns ``(p[3], objective, iterations)``."""
    goal = np.asarray(goal, dtype=np.float64)
    sf0 = float(np.sqrt(goal[0] * goal[0] + goal[1] * goal[1]))
    lo = np.array([-0.5, -0.5, sf0])
    hi = np.array([0.5, 0.5, np.inf])
    p = np.array([0.0, 0.0, sf0])
    f = float(objective(p, goal))
    lam = 1e-3
    wts = np.array([25.0, 25.0, 30.0])
    for it in range(max_iter):
        r, J = residual_jacobian(p, goal)
        grad = objective_grad(p, goal)
        H = 2.0 * (J.T * wts) @ J
        p1, p2, sf = p
        H[0, 0] += 648.0 * sf / 840.0
        H[1, 1] += 648.0 * sf / 840.0
        H[0, 1] += -81.0 * sf / 840.0
        H[1, 0] += -81.0 * sf / 840.0
        H[0, 2] += (648.0 * p1 - 81.0 * p2) / 840.0
        H[2, 0] += (648.0 * p1 - 81.0 * p2) / 840.0
        H[1, 2] += (648.0 * p2 - 81.0 * p1) / 840.0
        H[2, 1] += (648.0 * p2 - 81.0 * p1) / 840.0
        free = ~(((p <= lo) & (grad > 0)) | ((p >= hi) & (grad < 0)))
        pg = np.where(free, grad, 0.0)
        if np.abs(pg).max() <= tol * max(1.0, abs(f)):
            return p, f, it
        improved = False
        for _ in range(30):
            A = H + lam * np.diag(np.maximum(np.diag(H), 1e-12))
            idx = np.where(free)[0]
            step = np.zeros(3)
            try:
                step[idx] = -np.linalg.solve(A[np.ix_(idx, idx)], grad[idx])
            except np.linalg.LinAlgError:
                lam *= 10.0
                continue
            q = np.minimum(np.maximum(p + step, lo), hi)
            fq = float(objective(q, goal))
            if fq < f:
                p, f = q, fq
                lam = max(lam * 0.2, 1e-12)
                improved = True
                break
            lam *= 10.0
        if not improved:
            return p, f, it
    return p, f, max_iter
