"""GPU parity of the batched closed loop (Stanley + PID + steering filter around the RK4 step, K5) and of the
45-column DataLog it writes: against the literal reference's recorded frames and against the C oracle."""
import numpy as np
import pytest
import torch

from conftest import REL_TOL_F64, rel_err
from oracle import c_oracle, planar_numpy as pn
from python_motionplanning_b200 import TrackGains, VehicleParameters, datalog, workloads as wl
from python_motionplanning_b200.host_numerics import host_norm2_mode

pytestmark = pytest.mark.gpu
DT = 1e-4


@pytest.fixture(autouse=True, params=["auto", "closed_form"])
def friction_mode(request, engine):
    """Every test runs with the tabulated friction path of the fast kernels and with the closed form."""
    prev = engine.set_friction_mode(request.param)
    yield request.param
    engine.set_friction_mode(prev)


def _setup(engine):
    p = VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    engine.set_params(p)
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    return par


def _np(t):
    return t.cpu().numpy()


def test_tracking_frames_vs_literal_reference(engine, golden):
    """25 frames of the unmodified Car.drive loop, one vehicle per frame, each with that frame's waypoint list."""
    g = golden("tracking_frames.npz")
    _setup(engine)
    res = engine.track_closed_loop(g["start"].T, g["waypoints"], float(g["dt"]), 100, float(g["target_vel"]),
                                   ctrl0=g["ctrl0"][:, :3].T, wp_count=g["n_waypoints"], vehicles_per_set=1,
                                   store_stride=1, want_log=True, want_target_idx=True, norm_mode=int(g["norm2_mode"]))
    log = _np(res.log).transpose(2, 0, 1)
    ref = g["log"]
    assert np.array_equal(_np(res.target_idx).T, g["target_ids"])
    e = rel_err(log[:, :, 1:], ref[:, :, 1:])
    print(f"closed-loop frames vs literal: worst rel err {e.max():.3e}; per block state {e[:, :, :10].max():.2e} "
          f"sdot {e[:, :, 10:20].max():.2e} delta {e[:, :, 20].max():.2e} tau {e[:, :, 21:25].max():.2e} "
          f"outputs {e[:, :, 25:43].max():.2e} crosstrack {e[:, :, 43].max():.2e}")
    assert e.max() < REL_TOL_F64
    assert np.array_equal(log[:, :, 0], np.broadcast_to(np.arange(100) * float(g["dt"]), (len(ref), 100)))
    # one vehicle's rows in the reference's CSV form
    rows = datalog.vehicle_rows(res.log, 3)
    assert rows.shape == (100, 45) and np.array_equal(rows, log[3])


def test_tracking_fleet_vs_c_oracle(engine):
    """4,096 vehicles on 4 waypoint lists, 300 sub-steps (30 control updates): states, controls and indices."""
    par = _setup(engine)
    V, n_sets, N = 4096, 4, 300
    state0, wp = wl.tracking_fleet(V, n_sets)
    mode = host_norm2_mode()
    gains = c_oracle.track_gains()
    ctrl0 = np.zeros((3, V))
    ctrl0[2] = state0[0]
    ref = c_oracle.track_loop(state0, ctrl0, wp, None, par, DT, N, 25.0, gains, mode, vehicles_per_set=V // n_sets,
                              store_stride=10, want_log=True)
    res = engine.track_closed_loop(state0, wp, DT, N, 25.0, vehicles_per_set=V // n_sets, store_stride=10, want_log=True,
                                   want_target_idx=True, norm_mode=mode)
    tid = _np(res.target_idx)
    mism = int((tid != ref["target_idx"]).sum())
    e = rel_err(_np(res.log), ref["log"])
    print(f"fleet: worst rel err {e.max():.3e}; target-index mismatches {mism}/{tid.size}")
    assert mism == 0
    assert e.max() < REL_TOL_F64
    assert rel_err(_np(res.state_end), ref["state_end"]).max() < REL_TOL_F64
    assert rel_err(_np(res.ctrl_end), ref["ctrl_end"]).max() < REL_TOL_F64
    # no-log kernel (speculative RK4 step) gives the same trajectory
    res2 = engine.track_closed_loop(state0, wp, DT, N, 25.0, vehicles_per_set=V // n_sets, store_stride=10, norm_mode=mode)
    assert rel_err(_np(res2.traj), ref["traj"]).max() < REL_TOL_F64
    assert res2.log is None
    # resumable: 2 x 150 steps == 300 steps, bit for bit
    a = engine.track_closed_loop(state0, wp, DT, 150, 25.0, vehicles_per_set=V // n_sets, norm_mode=mode)
    b = engine.track_closed_loop(a.state_end, wp, DT, 150, 25.0, ctrl0=a.ctrl_end, vehicles_per_set=V // n_sets,
                                 step0=150, norm_mode=mode)
    assert torch.equal(b.state_end, res2.state_end) and torch.equal(b.ctrl_end, res2.ctrl_end)


def _single_update_indices(engine, wp, xs, ys, mode):
    """Target index of ONE control update per vehicle (one RK4 step follows, irrelevant here)."""
    V = len(xs)
    s0 = np.zeros((12, V))
    s0[0] = 20.0
    s0[3:7] = 20.0 / 0.308309813617345
    s0[8], s0[9] = xs, ys
    r = engine.track_closed_loop(s0, wp, DT, 1, 25.0, ctrl_every=1, want_target_idx=True, vehicles_per_set=V,
                                 store_stride=1, want_log=True, norm_mode=mode)
    return _np(r.target_idx)[0], _np(r.log)[0, 44]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_lookahead_index_exact_on_adversarial_waypoints(engine, mode):
    """The sub-linear nearest-waypoint search must return the reference's index (first strict minimum of the rounded
    norms, then the look-ahead walk) on lists built to break it: exact ties, duplicates, self-crossing loops, lengths
    around the 32-point chunk size, NaN points, far-away coordinates."""
    _setup(engine)
    rng = np.random.default_rng(5 + mode)
    gains = c_oracle.track_gains()
    cases = []
    t = np.linspace(0, 4 * np.pi, 1500)
    cases.append(np.stack([30 * np.sin(t), 20 * np.sin(2 * t)], 1))                       # figure eight, crosses itself
    cases.append(np.stack([np.linspace(-20, 20, 641), np.zeros(641)], 1))                 # symmetric: exact ties about x = 0
    cases.append(np.repeat(np.stack([np.linspace(0, 30, 100), np.linspace(0, 5, 100)], 1), 3, axis=0))   # duplicates
    cases.append(rng.uniform(-40, 40, (777, 2)))                                          # no spatial order at all
    cases.append(np.stack([np.arange(33) * 0.01, np.zeros(33)], 1))                       # one point past a chunk
    cases.append(np.array([[1.0, 2.0]]))                                                  # a single waypoint
    cases.append(np.stack([1.0e7 + np.arange(2000) * 0.01, -3.0e6 + np.zeros(2000)], 1))  # far from the origin
    nanwp = np.stack([np.arange(200) * 0.05, np.sin(np.arange(200) * 0.05)], 1)
    nanwp[[0, 31, 32, 64, 150], :] = np.nan
    cases.append(nanwp)
    cases.append(np.zeros((100, 2)))                                                      # all the same point
    # longer than one batch of coarse chunks (8 x 512 waypoints), ragged last chunks on every level, and a path that comes
    # back beside itself 2 cm away: chord culling must keep both branches alive
    u = np.arange(9003) * 0.01
    cases.append(np.stack([np.where(u < 45, u, 90 - u), np.where(u < 45, 0.0, 0.02) + 0.3 * np.sin(0.2 * np.minimum(u, 90 - u))], 1))
    sp = np.linspace(0.5, 12 * np.pi, 4611)
    cases.append(np.stack([0.4 * sp * np.cos(sp), 0.4 * sp * np.sin(sp)], 1))             # spiral: neighbouring turns 2.5 m apart
    for k, wp in enumerate(cases):
        V = 512
        lo, hi = np.nanmin(wp, 0) - 3.0, np.nanmax(wp, 0) + 3.0
        xs, ys = rng.uniform(lo[0], hi[0], V), rng.uniform(lo[1], hi[1], V)
        xs[:64], ys[:64] = wp[rng.integers(0, len(wp), 64)].T                             # exactly on waypoints
        if k == 1:
            xs[64:192] = 0.0                                                              # equidistant from +-x pairs
        if np.isnan(xs).any():
            bad = np.isnan(xs) | np.isnan(ys)
            xs[bad], ys[bad] = 0.5, 0.5
        idx, cte = _single_update_indices(engine, wp, xs, ys, mode)
        want = [c_oracle.stanley_control(wp, x, y, 0.0, 20.0, gains, mode) for x, y in zip(xs, ys)]
        assert np.array_equal(idx, [w[1] for w in want]), f"case {k}"
        assert np.array_equal(cte, [w[2] for w in want]), f"case {k}"                    # yaw = 0: cos/sin exact


def test_tracking_edge_cases(engine):
    _setup(engine)
    state0, wp = wl.tracking_fleet(128, 2, W=500)
    z = engine.track_closed_loop(state0[:, :0], wp, DT, 10)
    assert z.state_end.shape == (12, 0)
    r0 = engine.track_closed_loop(state0, wp, DT, 0)
    assert np.array_equal(_np(r0.state_end), state0) and np.array_equal(_np(r0.ctrl_end)[2], state0[0])
    with pytest.raises(ValueError):
        engine.track_closed_loop(state0, wp, DT, 10, step0=5)              # launches start on a control update
    with pytest.raises(ValueError):
        engine.track_closed_loop(state0, wp, DT, 10, vehicles_per_set=8)   # sets do not cover the fleet
    with pytest.raises(ValueError):
        engine.track_closed_loop(state0, wp, DT, 10, wp_count=[500, 501])  # a count beyond the list
    # custom gains / ctrl_every / ragged set lengths against the oracle
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    g = TrackGains(k=5.0, k_soft=2.0, max_steer=0.3, kp=500.0, ki=10.0, kd=0.01, lookahead=3.0, deadband=0.05, steer_filter=0.02)
    gv = c_oracle.track_gains(5.0, 2.0, 0.3, 500.0, 10.0, 0.01, 3.0, 0.05, 0.02)
    cnt = np.array([500, 137], dtype=np.int32)
    mode = host_norm2_mode()
    ctrl0 = np.stack([np.full(128, 0.01), np.full(128, -0.2), state0[0] - 0.1])
    ref = c_oracle.track_loop(state0, ctrl0, wp, cnt, par, DT, 57, 18.0, gv, mode, ctrl_every=4, vehicles_per_set=64,
                              store_stride=1, want_log=True)
    res = engine.track_closed_loop(state0, wp, DT, 57, 18.0, gains=g, ctrl0=ctrl0, wp_count=cnt, vehicles_per_set=64,
                                   ctrl_every=4, store_stride=1, want_log=True, want_target_idx=True, norm_mode=mode)
    assert np.array_equal(_np(res.target_idx), ref["target_idx"])
    assert rel_err(_np(res.log), ref["log"]).max() < REL_TOL_F64


def test_tracking_time_sliced_launch_equals_plain_launches(engine):
    """A launch of more than one wave of vehicles is cut into (block, time-chunk) items on persistent CTAs (the carried
    state, controller state and search hints go through L2): every output must equal, bit for bit, what plain launches
    (one wave each) produce -- trajectories, DataLog rows, look-ahead indices and resumable end state."""
    _setup(engine)
    n_sets, vps, N = 12, 4096, 200                      # 49,152 vehicles = 768 blocks: more than the 592 resident CTAs
    V = n_sets * vps
    st0, wps = wl.tracking_fleet(V, n_sets)
    mode = host_norm2_mode()
    for kw in (dict(store_stride=20, want_target_idx=True), dict(store_stride=20, want_log=True), dict()):
        big = engine.track_closed_loop(st0, wps, DT, N, 25.0, vehicles_per_set=vps, norm_mode=mode, **kw)
        parts = []
        for s0 in range(0, n_sets, 4):                  # 16,384 vehicles = 256 blocks per launch: the plain kernel
            lo, hi = s0 * vps, (s0 + 4) * vps
            parts.append(engine.track_closed_loop(st0[:, lo:hi], wps[s0:s0 + 4], DT, N, 25.0, vehicles_per_set=vps,
                                                  norm_mode=mode, **kw))
        assert torch.equal(big.state_end, torch.cat([p.state_end for p in parts], dim=1))
        assert torch.equal(big.ctrl_end, torch.cat([p.ctrl_end for p in parts], dim=1))
        if big.traj is not None:
            assert torch.equal(big.traj, torch.cat([p.traj for p in parts], dim=2))
        if big.log is not None:
            assert torch.equal(big.log, torch.cat([p.log for p in parts], dim=2))
        if big.target_idx is not None:
            assert torch.equal(big.target_idx, torch.cat([p.target_idx for p in parts], dim=1))
    # and against the C oracle on a strided subsample of the big launch (the oracle scans 3,000 waypoints per update)
    par = _setup(engine)
    big = engine.track_closed_loop(st0, wps, DT, N, 25.0, vehicles_per_set=vps, norm_mode=mode, want_target_idx=True)
    sub = np.arange(0, vps, 64)
    for s in (0, 5, 11):
        idx = s * vps + sub
        c0 = np.zeros((3, len(idx)))
        c0[2] = st0[0, idx]
        ref = c_oracle.track_loop(st0[:, idx], c0, wps[s:s + 1], None, par, DT, N, 25.0, c_oracle.track_gains(), mode,
                                  vehicles_per_set=len(idx))
        assert np.array_equal(_np(big.target_idx)[:, idx], ref["target_idx"])
        assert rel_err(_np(big.state_end)[:, idx], ref["state_end"]).max() < REL_TOL_F64


def test_tracking_sliced_launch_with_ragged_blocks(engine):
    """Time-sliced launch whose sets do not fill their last 128-thread block (1,000 vehicles per set = 7 full blocks + 104
    threads, i.e. a partly filled warp): the CTA-wide rendezvous in front of every control update counts arrivals from
    the threads that own no vehicle as well.  Results must equal plain launches of a few sets each, bit for bit."""
    _setup(engine)
    n_sets, vps, N = 40, 1000, 120                      # 320 blocks: more than the resident CTAs -> sliced
    V = n_sets * vps
    st0, wps = wl.tracking_fleet(V, n_sets, W=1500)
    mode = host_norm2_mode()
    big = engine.track_closed_loop(st0, wps, DT, N, 25.0, vehicles_per_set=vps, norm_mode=mode, want_target_idx=True,
                                   store_stride=40)
    parts = []
    for s0 in range(0, n_sets, 10):                     # 80 blocks per launch: one block per CTA, no ticket
        lo, hi = s0 * vps, (s0 + 10) * vps
        parts.append(engine.track_closed_loop(st0[:, lo:hi], wps[s0:s0 + 10], DT, N, 25.0, vehicles_per_set=vps,
                                              norm_mode=mode, want_target_idx=True, store_stride=40))
    assert torch.equal(big.state_end, torch.cat([p.state_end for p in parts], dim=1))
    assert torch.equal(big.ctrl_end, torch.cat([p.ctrl_end for p in parts], dim=1))
    assert torch.equal(big.traj, torch.cat([p.traj for p in parts], dim=2))
    assert torch.equal(big.target_idx, torch.cat([p.target_idx for p in parts], dim=1))
    # a fleet that does not fill its sets at all (V < n_sets * vehicles_per_set) against the oracle
    par = _setup(engine)
    Vs = 3 * 70 - 11
    st1, wp1 = wl.tracking_fleet(3 * 70, 3, W=400)
    st1 = st1[:, :Vs]
    c0 = np.zeros((3, Vs))
    c0[2] = st1[0]
    ref = c_oracle.track_loop(st1, c0, wp1, None, par, DT, 60, 25.0, c_oracle.track_gains(), mode, vehicles_per_set=70)
    res = engine.track_closed_loop(st1, wp1, DT, 60, 25.0, vehicles_per_set=70, norm_mode=mode, want_target_idx=True)
    assert np.array_equal(_np(res.target_idx), ref["target_idx"])
    assert rel_err(_np(res.state_end), ref["state_end"]).max() < REL_TOL_F64
