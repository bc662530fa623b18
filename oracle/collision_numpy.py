"""NumPy oracle for the circle-offset collision test and best-path selection.  TEST INFRASTRUCTURE ONLY.

Restates
  * ``CollisionChecker.collision_check``        reference ``libs/motionplanner/collision_checker.py:32-117``
  * ``CollisionChecker.select_best_path_index`` reference ``libs/motionplanner/collision_checker.py:134-203``

Third-party arithmetic the reference leans on (sources not under /root/reference):
  * ``scipy.spatial.distance.cdist`` (euclidean, float64; reference pin scipy>=1.7.3, here 1.18.1):
    bit-identical to ``sqrt(dx*dx + dy*dy)`` with separately rounded products;
  * ``np.linalg.norm`` of a 2-element list (numpy>=1.19.2, here 2.3.5): ``sqrt(x.dot(x))`` through the
    host BLAS ``ddot``; on the build host this equals ``sqrt(fma(v1, v1, v0*v0))``.  It is BLAS/CPU
    dependent, so :func:`select_best_path_index` calls ``np.linalg.norm`` itself (exact by construction
    on whatever host it runs) and :func:`probe_norm2_mode` tells which closed form the host follows.

Booleans and indices produced here must be bit-exact against the literal reference; pinned by
``tests/golden/collision_*.npz`` (see ``make_golden.py``) and the Appendix-C vectors KAT3/KAT4.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

NORM2_NOFMA = 0      # sqrt(fl(v0*v0) + fl(v1*v1))
NORM2_FMA_V1 = 1     # sqrt(fma(v1, v1, fl(v0*v0)))   <- OpenBLAS Haswell/SkylakeX ddot tail loop
NORM2_FMA_V0 = 2     # sqrt(fma(v0, v0, fl(v1*v1)))


def circle_centres(px, py, pyaw, offsets):
    """Centres ``[P, n_pts, n_circ]`` exactly as collision_checker.py:86-89 forms them.

    ``x + off*cos(yaw)`` is two roundings (no FMA); ``yaw_j`` is ``path[2][j]`` for ``j < len(path[0])``
    (the 49/50 off-by-one of SURVEY.md §8a6 lives in the caller's data, not here).
    """
    px = np.asarray(px, dtype=np.float64)
    py = np.asarray(py, dtype=np.float64)
    n = px.shape[-1]
    pyaw = np.asarray(pyaw, dtype=np.float64)[..., :n]
    off = np.asarray(offsets, dtype=np.float64)
    c, s = np.cos(pyaw), np.sin(pyaw)
    cx = px[..., None] + off * c[..., None]
    cy = py[..., None] + off * s[..., None]
    return cx, cy


def collision_check_batch(px, py, pyaw, obstacles, offsets, radii, want_clearance=False, chunk_bytes=64 << 20):
    """``free[P]`` (True = collision-free) for P paths; optionally ``min_clearance[P] = min(d - r)``.

    Follows collision_checker.py:66-113 without the early exit (which cannot change the boolean).
    """
    cx, cy = circle_centres(px, py, pyaw, offsets)          # [P, n, k]
    P = cx.shape[0]
    obs = np.asarray(obstacles, dtype=np.float64).reshape(-1, 2)
    M = obs.shape[0]
    rad = np.asarray(radii, dtype=np.float64)
    free = np.ones(P, dtype=bool)
    clear = np.full(P, np.inf)
    if M == 0:
        return (free, clear) if want_clearance else free
    per_path = cx[0].size * M * 8 * 3
    step = max(1, int(chunk_bytes // max(per_path, 1)))
    ox = obs[:, 0]
    oy = obs[:, 1]
    for lo in range(0, P, step):
        hi = min(P, lo + step)
        dx = ox[None, None, None, :] - cx[lo:hi, :, :, None]
        dy = oy[None, None, None, :] - cy[lo:hi, :, :, None]
        d = np.sqrt(dx * dx + dy * dy)                       # cdist, :101-103
        d = d - rad[None, None, :, None]                     # :104-105
        free[lo:hi] = ~np.any(d < 0, axis=(1, 2, 3))         # :106-107  (d == r is free)
        if want_clearance:
            clear[lo:hi] = d.min(axis=(1, 2, 3))
    return (free, clear) if want_clearance else free


def collision_check(path, obstacles, offsets, radii):
    """Single-path form with the reference's list-of-lists arguments; returns a Python bool."""
    if len(path[0]) == 0:
        return True
    f = collision_check_batch(np.asarray(path[0])[None], np.asarray(path[1])[None],
                              np.asarray(path[2])[None], obstacles, offsets, radii)
    return bool(f[0])


def select_best_path_index(end_x, end_y, free, goal_xy, weight):
    """Literal-order restatement of collision_checker.py:162-203 on path end points.

    score_i = norm([x_i-gx, y_i-gy]) + sum_{j colliding, ascending} weight*norm([x_i-x_j, y_i-y_j]);
    colliding i -> inf; first strict minimum wins; nothing free -> None.  ``np.linalg.norm`` is called
    on 2-element lists exactly as the reference does, so the host BLAS rounding is reproduced.
    """
    P = len(end_x)
    best_index, best_score = None, float("inf")
    coll = [j for j in range(P) if not free[j]]
    norm = np.linalg.norm
    for i in range(P):
        if free[i]:
            score = norm([end_x[i] - goal_xy[0], end_y[i] - goal_xy[1]])
            for j in coll:
                score += weight * norm([end_x[i] - end_x[j], end_y[i] - end_y[j]])
        else:
            score = float("inf")
        if score < best_score:
            best_score, best_index = score, i
    return best_index


def norm2_exact(v0: float, v1: float, mode: int) -> float:
    """Closed forms of ``np.linalg.norm([v0, v1])`` evaluated with exact rational arithmetic."""
    import math
    if mode == NORM2_NOFMA:
        q = v0 * v0 + v1 * v1
    elif mode == NORM2_FMA_V1:
        q = float(Fraction(v1) * Fraction(v1) + Fraction(v0 * v0))
    elif mode == NORM2_FMA_V0:
        q = float(Fraction(v0) * Fraction(v0) + Fraction(v1 * v1))
    else:
        raise ValueError(mode)
    return math.sqrt(q)


def probe_norm2_mode(n=400, seed=7):
    """Which closed form the host's ``np.linalg.norm([a, b])`` follows; ``None`` if none of them."""
    rng = np.random.default_rng(seed)
    v = rng.uniform(-50, 50, size=(n, 2))
    lit = [float(np.linalg.norm([a, b])) for a, b in v]
    for mode in (NORM2_FMA_V1, NORM2_NOFMA, NORM2_FMA_V0):
        if all(norm2_exact(float(a), float(b), mode) == l for (a, b), l in zip(v, lit)):
            return mode
    return None
