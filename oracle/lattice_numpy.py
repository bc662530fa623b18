"""NumPy restatement of the data producers either side of the collision test.  TEST INFRASTRUCTURE ONLY.

  * ``sample_spiral``   reference ``libs/motionplanner/path_optimizer.py:131-174`` (+ ``thetaf`` :109-117)
  * ``transform_paths`` reference ``libs/motionplanner/local_planner.py:424-470``
  * ``box``             reference ``libs/utils/env.py:93-127`` (+ ``strightline`` :69-73)

These are SURVEY.md §8(f) row N1 and the obstacle generator of config 3.  They keep the reference's
length quirk: 50 arc-length samples give 50 yaws but 49 (x, y) (cumulative trapezoid without an
initial value), and ``transform_paths`` keeps the first ``len(x)`` yaws, so ``yaw_j`` lags ``(x_j, y_j)``
by one sample.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np

N_SAMPLES = 50  # np.linspace default, path_optimizer.py:160


def sample_spiral(p):
    """``[x(49), y(49), yaw(50)]`` python lists for optimisation parameters ``p = [p1, p2, sf]``."""
    p = [0.0, p[0], p[1], 0.0, p[2]]
    a = p[0]
    b = -(11.0 * p[0] / 2.0 - 9.0 * p[1] + 9.0 * p[2] / 2.0 - p[3]) / p[4]
    c = (9.0 * p[0] - 45.0 * p[1] / 2.0 + 18.0 * p[2] - 9.0 * p[3] / 2.0) / p[4] ** 2
    d = -(9.0 * p[0] / 2.0 - 27.0 * p[1] / 2.0 + 27.0 * p[2] / 2.0 - 9.0 * p[3] / 2.0) / p[4] ** 3
    s = np.linspace(0.0, p[4])
    t = np.array(a) * s + (np.array(b) / 2) * s ** 2 + (np.array(c) / 3) * s ** 3 + (np.array(d) / 4) * s ** 4
    ds = np.diff(s)
    ct, st = np.cos(t), np.sin(t)
    x = np.cumsum(ds * (ct[1:] + ct[:-1]) / 2.0)     # scipy cumulative_trapezoid, initial=None
    y = np.cumsum(ds * (st[1:] + st[:-1]) / 2.0)
    return [x.tolist(), y.tolist(), t.tolist()]


def transform_paths(paths, ego_state):
    """Ego-frame -> global-frame; iterates ``len(path[0])`` so the 50th yaw is dropped."""
    out = []
    ex, ey, eyaw = ego_state[0], ego_state[1], ego_state[2]
    for path in paths:
        n = len(path[0])
        xs = [ex + path[0][i] * cos(eyaw) - path[1][i] * sin(eyaw) for i in range(n)]
        ys = [ey + path[0][i] * sin(eyaw) + path[1][i] * cos(eyaw) for i in range(n)]
        ts = [path[2][i] + eyaw for i in range(n)]
        out.append([xs, ys, ts])
    return out


def _straight(x_start, x_end, y, ds):
    X = [x for x in np.arange(x_start, x_end, ds)]
    return X, [y] * len(X)


def box(corner, width, length, ds):
    """Obstacle outline points ``(X, Y)`` of one box, env.py:93-127 (four edges, same order)."""
    xc1, yc1 = corner
    xc2, yc2 = xc1 + length, yc1 + width
    X, Y = [], []
    x1, y1 = _straight(xc1, xc2, yc1, ds)
    X += x1
    Y += y1
    y2 = [y for y in np.arange(yc1, yc2, ds)]
    X += [xc2] * len(y2)
    Y += y2
    x3, y3 = _straight(xc2, xc1, yc2, -ds)
    X += x3
    Y += y3
    y4 = [y for y in np.arange(yc1, yc2, ds)]
    X += [xc1] * len(y4)
    Y += y4
    return X, Y
