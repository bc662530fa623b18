"""The reference's 45-column ``DataLog`` (drive.py:44, 145-151) and its CSV form (plots.py:17-36).

``Engine.track_closed_loop(..., want_log=True)`` (and ``rollout(..., want_aux=True)`` through
``assemble_open_loop``) produce the rows on the GPU; this module only names the columns, slices one
vehicle's rows out of the ``[n_out, 45, V]`` device layout and writes ``results/Results.csv`` in the byte
format ``data_cleaning`` produces (``pandas.DataFrame.to_csv``: unnamed index column, ``repr`` floats), so
the reference's ``plot_results`` works on GPU-generated data unchanged.
"""
from __future__ import annotations

import os

import numpy as np

# plots.py:19-27
DATALOG_COLUMNS = ['time',
                   'U', 'V', 'wz', 'wFL', 'wFR', 'wRL', 'wRR', 'yaw', 'x', 'y',
                   'U_dot', 'V_dot', 'wz_dot', 'wFL_dot', 'wFR_dot', 'wRL_dot',
                   'wRR_dot', 'yaw_dot', 'x_dot', 'y_dot',
                   'delta', 'tau_FL', 'tau_FR', 'tau_RL', 'tau_RR',
                   'Fx_FL', 'Fx_FR', 'Fx_RL', 'Fx_RR',
                   'Fy_FL', 'Fy_FR', 'Fy_RL', 'Fy_RR',
                   'Fz_FL', 'Fz_FR', 'Fz_RL', 'Fz_RR',
                   'sFL', 'sFR', 'sRL', 'sRR', 'Fxt_FL', 'Fyt_FL', 'crosstrack']
N_LOG = len(DATALOG_COLUMNS)
assert N_LOG == 45


def vehicle_rows(log, vehicle: int = 0) -> np.ndarray:
    """Rows ``[n_out, 45]`` of one vehicle from the device layout ``[n_out, 45, V]`` (tensor or ndarray)."""
    if hasattr(log, "detach"):
        log = log[:, :, vehicle].detach().cpu().numpy()
    else:
        log = np.asarray(log)[:, :, vehicle]
    return np.ascontiguousarray(log, dtype=np.float64)


def assemble_open_loop(traj, aux, delta, torque, dt: float, hold: int = 1, store_stride: int = 1, step0: int = 0,
                       crosstrack=np.nan) -> np.ndarray:
    """DataLog rows ``[n_out, 45, B]`` for OPEN-loop rollouts from ``RolloutResult.traj`` / ``.aux`` and the
    control segments (front-steer / equal-torque layout): what drive.py:145-151 would have logged had the
    recorded controls been applied.  ``crosstrack`` has no open-loop meaning and defaults to NaN."""
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    traj, aux, delta, torque = (to_np(a) for a in (traj, aux, delta, torque))
    n_out, _, B = traj.shape
    rows = np.empty((n_out, N_LOG, B))
    steps = step0 + (np.arange(n_out) + 1) * store_stride - 1        # the sub-step each stored row belongs to
    rows[:, 0, :] = (steps * dt)[:, None]
    rows[:, 1:11, :] = traj
    rows[:, 11:21, :] = aux[:, :10, :]
    seg = steps // hold
    d = np.broadcast_to(delta[seg, 0, :], (n_out, B))
    t = torque[seg, :, :]
    rows[:, 21, :] = d
    rows[:, 22:26, :] = t if t.shape[1] == 4 else np.broadcast_to(t, (n_out, 4, B))
    rows[:, 26:44, :] = aux[:, 10:, :]
    rows[:, 44, :] = crosstrack
    return rows


def clean(DataLog: np.ndarray) -> np.ndarray:
    """``data_cleaning``'s row filter: rows that are entirely zero are dropped (plots.py:18)."""
    DataLog = np.asarray(DataLog, dtype=np.float64)
    return DataLog[~np.all(DataLog == 0, axis=1)]


def write_results_csv(DataLog: np.ndarray, path: str = os.path.join("results", "Results.csv")) -> str:
    """Write ``DataLog`` rows as the reference's ``results/Results.csv`` (plots.py:17-36): header
    ``,time,U,...,crosstrack``, one line per non-zero row, first field the row number, floats as ``repr``."""
    rows = clean(DataLog)
    if rows.ndim != 2 or rows.shape[1] != N_LOG:
        raise ValueError(f"DataLog must be [n, {N_LOG}]")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)       # the reference fails when results/ is missing (plots.py:28-33)
    with open(path, "w", newline="") as f:
        f.write("," + ",".join(DATALOG_COLUMNS) + "\n")
        for i, r in enumerate(rows):
            f.write(str(i) + "," + ",".join(_fmt(v) for v in r) + "\n")
    return path


def _fmt(v: float) -> str:
    # pandas writes float64 through repr(); NaN as the empty string (na_rep default), inf as "inf"
    if v != v:
        return ""
    return repr(float(v))
