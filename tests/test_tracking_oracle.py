"""CPU: the closed-loop tracking oracle (oracle.c: Stanley + PID + steering filter around the RK4 step) is held to
the literal reference -- 25 frames of the unmodified ``Car.drive`` loop (tests/golden/tracking_frames.npz, made by
oracle/make_golden.py --only tracking) -- and the DataLog CSV writer to pandas' ``to_csv`` format."""
import io
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import c_oracle, planar_numpy as pn
from python_motionplanning_b200 import datalog


def _par():
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    return par


def test_tracking_oracle_vs_literal_frames(golden):
    g = golden("tracking_frames.npz")
    gains = c_oracle.track_gains()
    assert np.array_equal(gains[:6], g["gains"]) and gains[6] == g["lookahead"] and gains[7] == g["deadband"]
    assert gains[8] == g["steer_filter"]
    res = c_oracle.track_loop(g["start"].T, g["ctrl0"][:, :3].T, g["waypoints"], g["n_waypoints"], _par(), float(g["dt"]),
                              100, float(g["target_vel"]), gains, int(g["norm2_mode"]), vehicles_per_set=1,
                              store_stride=1, want_log=True)
    log = res["log"].transpose(2, 0, 1)                       # [frame][sub-step][45]
    ref = g["log"]
    assert np.array_equal(res["target_idx"].T, g["target_ids"])          # look-ahead indices: exact
    assert np.array_equal(log[:, :, 21], ref[:, :, 21])                  # filtered steering angle: exact
    assert np.array_equal(log[:, :, 22:26], ref[:, :, 22:26])            # PID torque: exact
    assert np.array_equal(log[:, :, 44], ref[:, :, 44])                  # crosstrack error: exact
    assert rel_err(log[:, :, 1:], ref[:, :, 1:]).max() < 1e-12           # states, derivatives, outputs
    # the time column is (frame*100 + i) * dt (drive.py:145); the oracle starts every frame at step0 = 0
    t = (g["frame_index"][:, None] * 100 + np.arange(100)[None]) * float(g["dt"])
    assert np.array_equal(ref[:, :, 0], t)
    assert np.array_equal(log[:, :, 0], np.broadcast_to(np.arange(100) * float(g["dt"]), (len(t), 100)))


def test_stanley_single_calls_vs_literal(golden):
    """Each control update of the recorded frames, one call at a time, from the logged state."""
    g = golden("tracking_frames.npz")
    gains = c_oracle.track_gains()
    mode = int(g["norm2_mode"])
    for f in range(0, len(g["log"]), 6):
        wp = g["waypoints"][f, :g["n_waypoints"][f]]
        for u in range(1, 10):                                 # state before sub-step 10*u = row 10*u - 1
            row = g["log"][f, 10 * u - 1]
            x, y, yaw, v = row[9], row[10], row[8], row[1]
            _, idx, cte = c_oracle.stanley_control(wp, x, y, yaw, v, gains, mode)
            assert idx == g["target_ids"][f, u]
            assert cte == g["log"][f, 10 * u, 44]


def test_results_csv_matches_pandas(tmp_path, golden):
    pd = pytest.importorskip("pandas")
    g = golden("tracking_frames.npz")
    rows = g["log"][:2].reshape(-1, 45).copy()
    rows[5] = 0.0                                              # an all-zero row is dropped (plots.py:18)
    rows[7, 44] = np.nan
    rows[9, 3] = -0.0
    rows[11, 2] = 1e-300
    rows[12, 4] = np.inf
    out = datalog.write_results_csv(rows, str(tmp_path / "results" / "Results.csv"))
    want = io.StringIO()
    pd.DataFrame(rows[~np.all(rows == 0, axis=1)], columns=datalog.DATALOG_COLUMNS).to_csv(want, lineterminator="\n")
    assert open(out).read() == want.getvalue()
    back = pd.read_csv(out, index_col=0)
    assert list(back.columns) == datalog.DATALOG_COLUMNS and len(back) == len(rows) - 1


def test_open_loop_datalog_assembly():
    """assemble_open_loop lays traj / aux / controls out as drive.py:145-151 does."""
    from python_motionplanning_b200 import workloads as wl
    B, N = 8, 40
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    ref = c_oracle.rollout(s0, d, t, _par(), wl.DT, N, hold=wl.HOLD, store_stride=1, want_aux=True)
    rows = datalog.assemble_open_loop(ref["traj"], ref["aux"], d, t, wl.DT, hold=wl.HOLD)
    assert rows.shape == (N, 45, B)
    n = 23
    assert rows[n, 0, 3] == n * wl.DT
    assert np.array_equal(rows[n, 1:11, 3], ref["traj"][n, :, 3]) and np.array_equal(rows[n, 11:21, 3], ref["aux"][n, :10, 3])
    assert rows[n, 21, 3] == d[n // wl.HOLD, 0, 3] and np.all(rows[n, 22:26, 3] == t[n // wl.HOLD, 0, 3])
    assert np.array_equal(rows[n, 26:44, 3], ref["aux"][n, 10:, 3]) and np.isnan(rows[n, 44, 3])
