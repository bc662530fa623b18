"""Drop-in ``CollisionChecker`` with the reference's call surface, computed on the GPU.

Mirrors ``libs/motionplanner/collision_checker.py`` of the reference:
  * ``CollisionChecker(circle_offsets, circle_radii, weight)``                         (:16-20)
  * ``.collision_check(path, obstacles) -> bool``   one path per call, True = free       (:32-117)
  * ``.select_best_path_index(paths, collision_check_array, goal_state) -> int | None``  (:134-203)
plus the batch form the planner's fan-out maps onto:
  * ``.collision_check_paths(paths, obstacles) -> list[bool]``   all paths in ONE kernel launch
Inputs are the reference's nested Python lists (or ndarrays).  Reference quirks kept: only the first
``len(path[0])`` yaws are read (the 49/50 off-by-one); ``dist == r`` is free; an empty
obstacle list or empty path is free; ties pick the lowest index; nothing free returns ``None``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine
from .vehicle_model import default_engine


class CollisionChecker:
    def __init__(self, circle_offsets, circle_radii, weight, engine: Optional[Engine] = None):
        self._circle_offsets = circle_offsets
        self._circle_radii = circle_radii
        self._weight = weight
        self._engine = engine

    def _eng(self) -> Engine:
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    # ------------------------------------------------------------------ batch
    def collision_check_paths(self, paths, obstacles):
        """Collision flags for a list of paths ``[x_points, y_points, t_points]`` (ragged lengths allowed)."""
        n_paths = len(paths)
        if n_paths == 0:
            return []
        lens = [len(p[0]) for p in paths]
        n = max(lens)
        obs = np.asarray(obstacles, dtype=np.float64).reshape(-1, 2)
        if n == 0 or obs.shape[0] == 0:
            return [True] * n_paths
        px = np.empty((n_paths, n))
        py = np.empty((n_paths, n))
        pt = np.empty((n_paths, n))
        for i, p in enumerate(paths):
            m = lens[i]
            if m == 0:
                # an empty path is free; park its points where no obstacle can reach
                px[i], py[i], pt[i] = np.inf, np.inf, 0.0
                continue
            px[i, :m], py[i, :m], pt[i, :m] = p[0], p[1], p[2][:m]
            # ragged: repeat the last point (testing a point twice cannot change the verdict)
            px[i, m:], py[i, m:], pt[i, m:] = px[i, m - 1], py[i, m - 1], pt[i, m - 1]
        free = self._eng().collision_check_batch(px, py, pt, obs, list(self._circle_offsets), list(self._circle_radii))
        return [bool(f) for f in free.cpu().numpy()]

    # ------------------------------------------------- reference signatures
    def collision_check(self, paths, obstacles):
        """Single path (the reference treats its ``paths`` argument as ONE path, :63) -> bool."""
        return self.collision_check_paths([paths], obstacles)[0]

    def select_best_path_index(self, paths, collision_check_array, goal_state):
        n_paths = len(paths)
        if n_paths == 0:
            return None
        ex = np.array([p[0][-1] for p in paths], dtype=np.float64)
        ey = np.array([p[1][-1] for p in paths], dtype=np.float64)
        free = np.array([bool(collision_check_array[i]) for i in range(n_paths)], dtype=np.uint8)
        return self._eng().select_best_path_index_batch(ex, ey, free, goal_state[:2], float(self._weight))
