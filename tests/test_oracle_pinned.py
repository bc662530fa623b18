"""Pin the oracle (NumPy + plain-C restatements) to outputs of the literal reference.

The reference ships no tests or golden vectors; ``tests/golden/*.npz`` were produced by
``oracle/make_golden.py`` executing the unmodified reference in the build container.  These tests need
no GPU and no reference checkout.
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import c_oracle, collision_numpy as cn, planar_numpy as pn
from python_motionplanning_b200 import workloads as wl

P = pn.VehicleParams()


def test_vehicle_params_derived(golden):
    g = golden("planar_model.npz")
    for k in ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw", "BFL", "CFL", "DFL"):
        assert getattr(P, k) == float(g["param_" + k]), k
    # Appendix C / §8a1 probed values
    assert P.rw == 0.308309813617345 and P.b == 1.5708108108108108 and P.Izz == 1948.2304506781593


def test_planar_model_numpy_vs_literal(golden):
    g = golden("planar_model.npz")
    n = len(g["states"])
    sd, vx, vy, ax, ay, out, axc, ayc = pn.planar_model(g["states"].T, g["torque"].T, g["mu_max"].T, g["delta"].T, P,
                                                       g["ax_prev"], g["ay_prev"])
    # bitwise except the scalar reference's pow(x, 2) (1 ulp in 0.08 % of inputs)
    assert rel_err(sd.T, g["state_dot"], 1e-300).max() < 1e-12
    assert rel_err(out.T, g["outputs"], 1e-300).max() < 1e-12
    misc = np.stack([vx, vy, ax, ay, axc, ayc], axis=1)
    assert rel_err(misc, g["misc"], 1e-300).max() < 1e-11
    assert np.mean(sd.T == g["state_dot"]) > 0.95
    assert n >= 64


def test_planar_model_kat1_hex(golden):
    """literal values quoted as hex floats."""
    g = golden("planar_model.npz")
    want = [float.fromhex(h) for h in ("0x1.bf1b72b75c814p+1", "-0x1.dc8697397ae5ap+1", "0x1.affe5b70fc6f7p+0",
                                       "-0x1.d72f42a883cb7p+9", "-0x1.118fc36753866p+9", "-0x1.faedb92ad939cp+7",
                                       "0x1.ff6e153a7f1dcp+4", "0x1.999999999999ap-3", "0x1.2f57f0971f0a0p+4",
                                       "0x1.98d62d86c595cp+2")]
    assert list(g["state_dot"][0]) == want
    sd = pn.planar_model(g["states"][0], g["torque"][0], g["mu_max"][0], g["delta"][0], P, 0.4, -0.7)[0]
    assert rel_err(sd, want, 1e-300).max() < 1e-15
    # zero-slip state: exactly zero derivatives except x_dot = U
    assert list(g["state_dot"][1]) == [0, 0, 0, 0, 0, 0, 0, 0, 25, 0]
    sd0 = pn.planar_model(g["states"][1], g["torque"][1], g["mu_max"][1], g["delta"][1], P, 0.0, 0.0)[0]
    assert list(sd0) == [0, 0, 0, 0, 0, 0, 0, 0, 25, 0]


def test_planar_model_c_vs_literal(golden):
    g = golden("planar_model.npz")
    for i in range(len(g["states"])):
        sd, misc, out = c_oracle.planar_model(g["states"][i], g["torque"][i], g["mu_max"][i], g["delta"][i], P,
                                              g["ax_prev"][i], g["ay_prev"][i])
        assert rel_err(sd, g["state_dot"][i], 1e-300).max() < 1e-12
        assert rel_err(out, g["outputs"][i], 1e-300).max() < 1e-12
        assert rel_err(misc, g["misc"][i], 1e-300).max() < 1e-11


def test_rk4_step_vs_literal(golden):
    g = golden("planar_model.npz")
    st, sd, out, axc, ayc = pn.planar_model_rk4(g["states"].T, g["torque"].T, g["mu_max"].T, g["delta"].T, P,
                                                g["ax_prev"], g["ay_prev"], float(g["dt"]))
    assert rel_err(st.T, g["rk4_state"], 1e-300).max() < 1e-14
    assert rel_err(sd.T, g["rk4_state_dot"], 1e-300).max() < 1e-11
    assert rel_err(out.T, g["rk4_outputs"], 1e-300).max() < 1e-11
    assert rel_err(np.stack([axc, ayc], 1), g["rk4_axay"], 1e-300).max() < 1e-10


def test_rollout_cfg2_subsample_vs_literal(golden):
    """256 rollouts x 500 steps of config 2: literal reference vs NumPy and C oracles."""
    g = golden("rollout_cfg2_sub.npz")
    s0, d, t = g["state0"], g["delta"], g["torque"]
    steps = list(g["check_steps"])
    # the fixture's inputs are the seeded config-2 batch at the stored indices
    s0f, df, tf = wl.config2_rollouts(B=65536, n_steps=500)
    assert np.array_equal(s0f[:, g["index"]], s0) and np.array_equal(df[:, :, g["index"]], d)
    res = pn.rollout(s0[:10], d, t, P, float(g["dt"]), 500, hold=int(g["hold"]))
    par = c_oracle.make_params(P)
    par[0].D[:] = (1.0,) * 4
    resc = c_oracle.rollout(s0, d, t, par, float(g["dt"]), 500, hold=int(g["hold"]), store_stride=1)
    for k, n in enumerate(steps):
        lit = g["states"][:, k, :10].T
        assert rel_err(res["traj"][n - 1], lit).max() < 1e-12, n
        assert rel_err(resc["traj"][n - 1], lit).max() < 1e-12, n
    lit_axay = g["states"][:, -1, 10:].T
    assert np.abs(res["ax_end"] - lit_axay[0]).max() < 1e-9
    assert np.abs(resc["state_end"][10:] - lit_axay).max() < 1e-9


def test_kat2_rollout():
    """SURVEY.md Appendix C KAT2 (literal reference, chained ax_prev/ay_prev)."""
    rw = P.rw
    st = np.array([25, 0, 0, 25 / rw, 25 / rw, 25 / rw, 25 / rw, 0.1, 1.0, 2.0])[:, None]
    res = pn.rollout(st, np.full((1, 1, 1), 0.05), np.full((1, 1, 1), 100.0), P, 1e-4, 500, hold=500)
    s1 = [24.999988855398556, 0.0004945442201748125, 0.0006301193542856004, 81.09358272887985, 81.0936370444646,
          81.09698421572764, 81.09706056433605, 0.10000003151355268, 1.0024875073659343, 2.0002496081255767]
    s500 = [25.02201847903116, 0.10174772160255088, 0.22727899273014812, 80.97629013263781, 81.83751630278395,
            80.95194876798494, 81.85810509344246, 0.10633896619693778, 2.2436021835767246, 2.131184130686531]
    assert rel_err(res["traj"][0][:, 0], s1, 1e-300).max() < 1e-13
    assert rel_err(res["traj"][499][:, 0], s500, 1e-300).max() < 1e-12
    assert abs(res["ax_end"][0] - 0.4780425299128667) < 1e-11 and abs(res["ay_end"][0] - 5.547717644536824) < 1e-11


def test_closed_loop_replay_vs_literal(golden):
    """Config 1: replay the recorded (delta, torque) of the reference's closed loop, 40,000 steps."""
    g = golden("closedloop_cfg1.npz")
    n_ctrl = len(g["delta"])
    par = c_oracle.make_params(P)
    par[0].D[:] = (1.0,) * 4
    res = c_oracle.rollout(g["state0"][:, None], g["delta"][:, None, None], g["torque"][:, None, None], par,
                           float(g["dt"]), n_ctrl * 10, hold=10, store_stride=10, nthreads=1)
    assert rel_err(res["traj"][:, :, 0], g["state_every10"]).max() < 1e-10
    # first frame, every sub-step, NumPy oracle incl. the logged state_dot / outputs (drive.py:147-150)
    r2 = pn.rollout(g["state0"][:10, None], g["delta"][:10, None, None], g["torque"][:10, None, None], P, float(g["dt"]),
                    100, hold=10, want_aux=True)
    assert rel_err(r2["traj"][:, :, 0], g["first_frame_states"]).max() < 1e-13
    assert rel_err(r2["state_dot"][:, :, 0], g["first_frame_sdot"]).max() < 1e-9
    assert rel_err(r2["outputs"][:, :, 0], g["first_frame_outputs"]).max() < 1e-9


# ------------------------------------------------------------------------------------------ collision
def test_collision_subsample_bit_exact(golden):
    g = golden("collision_cfg3_sub.npz")
    w = wl.config3_lattice()
    assert np.array_equal(w["px"][g["index"]], g["px"]) and np.array_equal(w["obstacles"], g["obstacles"])
    f_np = cn.collision_check_batch(g["px"], g["py"], g["pyaw"], g["obstacles"], g["offsets"], g["radii"])
    f_c, _, _ = c_oracle.collision_check(g["px"], g["py"], g["pyaw"], g["obstacles"], g["offsets"], g["radii"])
    f_c2, clr, _ = c_oracle.collision_check(g["px"], g["py"], g["pyaw"], g["obstacles"], g["offsets"], g["radii"],
                                            want_clearance=True)
    assert np.array_equal(f_np, g["free"]) and np.array_equal(f_c, g["free"]) and np.array_equal(f_c2, g["free"])
    assert np.array_equal(clr >= 0, g["free"])
    assert 0.2 <= g["free"].mean() <= 0.8


def test_collision_kat3(golden):
    g = golden("collision_cfg3_sub.npz")
    path = [[float(i) for i in range(1, 50)], [0.0] * 49, [0.0] * 50]
    assert list(g["kat3_free"]) == [True, False, True, False, True, True]      # Appendix C KAT3
    for obs, cnt, want in zip(g["kat3_obstacles"], g["kat3_counts"], g["kat3_free"]):
        o = obs[:cnt]
        assert cn.collision_check(path, o, g["offsets"], g["radii"]) == bool(want)
        f, _, _ = c_oracle.collision_check(np.array([path[0]]), np.array([path[1]]), np.array([path[2]]), o,
                                           g["offsets"], g["radii"])
        assert bool(f[0]) == bool(want)


def test_collision_boundary_ulps(golden):
    """Obstacle points within a few ulps of the circle boundary: pins no-FMA cdist + strict `<`."""
    g = golden("collision_cfg3_sub.npz")
    n = len(g["bnd_x"])
    free_np = np.array([cn.collision_check_batch(g["bnd_x"][i:i + 1, None], g["bnd_y"][i:i + 1, None],
                                                 g["bnd_yaw"][i:i + 1, None], [[g["bnd_ox"][i], g["bnd_oy"][i]]],
                                                 g["offsets"], g["radii"])[0] for i in range(n)])
    assert np.array_equal(free_np, g["bnd_free"])
    free_c = np.array([c_oracle.collision_check(g["bnd_x"][i:i + 1, None], g["bnd_y"][i:i + 1, None],
                                                g["bnd_yaw"][i:i + 1, None], [[g["bnd_ox"][i], g["bnd_oy"][i]]],
                                                g["offsets"], g["radii"])[0][0] for i in range(n)])
    assert np.array_equal(free_c, g["bnd_free"])
    assert 0.1 < g["bnd_free"].mean() < 0.9


def test_select_best_vs_literal(golden):
    g = golden("collision_cfg3_sub.npz")
    w = wl.config3_lattice()
    ex, ey = w["px"][:, -1], w["py"][:, -1]
    free = g["sel_free512"]
    mode = cn.probe_norm2_mode()
    for lo, n, best in g["sel_cases"]:
        want = None if best < 0 else int(best)
        assert cn.select_best_path_index(ex[lo:lo + n], ey[lo:lo + n], free[lo:lo + n], g["goal"], float(g["weight"])) == want
        if mode is not None and mode == int(g["norm2_mode"]):
            got, _ = c_oracle.select_best(ex[lo:lo + n], ey[lo:lo + n], free[lo:lo + n], g["goal"], float(g["weight"]), mode)
            assert got == want


def test_select_best_kat4(golden):
    g = golden("collision_cfg3_sub.npz")
    assert list(g["kat4_best"]) == [1, 2, -1, 0]                                 # Appendix C KAT4
    ex, ey = np.full(5, 49.0), np.array([-4.0, -2.0, 0.0, 2.0, 4.0])
    for flags, best in zip(g["kat4_flags"], g["kat4_best"]):
        want = None if best < 0 else int(best)
        assert cn.select_best_path_index(ex, ey, flags, [49, 0], 10) == want
        for mode in (0, 1, 2):     # exact ties: every rounding mode agrees
            assert c_oracle.select_best(ex, ey, flags, [49, 0], 10, mode)[0] == want


def test_norm2_closed_form_matches_host():
    """The C closed forms reproduce this host's np.linalg.norm([a, b]) bit for bit in the probed mode."""
    mode = cn.probe_norm2_mode()
    assert mode is not None, "host np.linalg.norm follows none of the known closed forms"
    rng = np.random.default_rng(5)
    v = rng.uniform(-80, 80, (2000, 2))
    lit = np.array([np.linalg.norm([a, b]) for a, b in v])
    assert np.array_equal(c_oracle.norm2(v[:, 0], v[:, 1], mode), lit)


def test_closed_loop_planner_flags_vs_literal(golden):
    """Config 1: the 7-path lattices the reference planned, its collision flags and chosen indices."""
    g = golden("closedloop_cfg1.npz")
    obs = g["obstacle_xy"]
    assert obs.shape == (106, 2)
    frames = g["plan_path_frames"]
    paths = g["plan_paths"]            # [n, 7, 3, 49]
    for k, f in enumerate(frames):
        px, py, pyaw = paths[k, :, 0], paths[k, :, 1], paths[k, :, 2]
        free = cn.collision_check_batch(px, py, pyaw, obs, wl.CIRCLE_OFFSETS, wl.CIRCLE_RADII)
        assert np.array_equal(free, g["plan_flags"][f]), f
        best = cn.select_best_path_index(px[:, -1], py[:, -1], free, g["plan_goal"][f], wl.PATH_SELECT_WEIGHT)
        assert (-1 if best is None else best) == g["plan_best"][f], f


def test_planner_core_with_dropped_paths_vs_literal(golden):
    """plan_invalid.npz (literal plan_paths -> transform_paths -> collision_check -> select_best_path_index with
    unreachable goals): the oracle pipeline applied to the FILTERED list reproduces flags and index, and treating a
    dropped path as "excluded" in the full list (what the device does) picks the same path."""
    g = golden("plan_invalid.npz")
    mode = cn.probe_norm2_mode()
    for c in range(int(g["n_cases"])):
        validity, flags, want = g[f"c{c}_validity"], g[f"c{c}_flags"], int(g[f"c{c}_best"])
        ends, goal = g[f"c{c}_ends"], g[f"c{c}_goal_state"]
        assert len(flags) == int(validity.sum()) == len(ends)
        if len(ends) == 0:
            assert want == -1
            continue
        if mode is not None and mode == int(g["norm2_mode"]):
            got, _ = c_oracle.select_best(ends[:, 0], ends[:, 1], flags, goal[:2], float(g["weight"]), mode)
            assert (-1 if got is None else got) == want, c
        got_np = cn.select_best_path_index(ends[:, 0], ends[:, 1], flags, goal[:2], float(g["weight"]))
        assert (-1 if got_np is None else got_np) == want, c
