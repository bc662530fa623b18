"""Host glue either side of the device lattice pipeline (``Engine.plan_lattice``): the planner's goal-state set.

``goal_state_set`` mirrors ``LocalPlanner.get_goal_state_set`` (reference libs/motionplanner/local_planner.py:154-275):
the goal waypoint is expressed in the vehicle frame and ``num_paths`` goal states are laid out perpendicular to the
path heading, ``path_offset`` apart.  It is a few dozen scalar operations per plan and stays on the host; its output
(transposed to ``[3, P]``) is the input of ``Engine.optimize_spirals`` / ``Engine.plan_lattice``.
"""
from __future__ import annotations

from math import cos, pi, sin

import numpy as np


def goal_state_set(goal_index: int, goal_state, waypoints, ego_state, num_paths: int = 7, path_offset: float = 2.0):
    """``[[x, y, t, v]] * num_paths`` in the vehicle frame; same arithmetic and operator order as the reference."""
    if goal_index < len(waypoints) - 1:                                     # :209-216
        delta_x = waypoints[goal_index + 1][0] - waypoints[goal_index][0]
        delta_y = waypoints[goal_index + 1][1] - waypoints[goal_index][1]
    else:
        delta_x = waypoints[goal_index][0] - waypoints[goal_index - 1][0]
        delta_y = waypoints[goal_index][1] - waypoints[goal_index - 1][1]
    heading = np.arctan2(delta_y, delta_x)
    gx = goal_state[0] - ego_state[0]                                       # :223-226
    gy = goal_state[1] - ego_state[1]
    theta = -ego_state[2]                                                   # :236-239
    goal_x = gx * cos(theta) - gy * sin(theta)
    goal_y = gx * sin(theta) + gy * cos(theta)
    goal_t = heading - ego_state[2]                                         # :246
    goal_v = goal_state[2]
    if goal_t > pi:                                                         # :252-255
        goal_t -= 2 * pi
    elif goal_t < -pi:
        goal_t += 2 * pi
    out = []
    for i in range(num_paths):                                              # :258-273
        offset = (i - num_paths // 2) * path_offset
        x_offset = offset * np.cos(goal_t + pi / 2)
        y_offset = offset * np.sin(goal_t + pi / 2)
        out.append([goal_x + x_offset, goal_y + y_offset, goal_t, goal_v])
    return out
