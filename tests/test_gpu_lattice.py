"""GPU parity of device-side lattice generation (K6: sample_spiral + transform_paths) and of the fully
device-resident lattice pipeline (spiral parameters -> paths -> collision flags -> best index)."""
import numpy as np
import pytest

from oracle import c_oracle
from python_motionplanning_b200 import workloads as wl
from python_motionplanning_b200.host_numerics import host_norm2_mode

pytestmark = pytest.mark.gpu
OFF, RAD, W = list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT


def _np(t):
    return t.cpu().numpy()


def _close(a, ref, tol=1e-12):
    return (np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)).max() < tol


def test_lattice_vs_literal_reference(engine, golden):
    g = golden("lattice_paths.npz")
    ego = g["ego"].T.copy()
    r = engine.sample_lattice(g["kappa1"], g["kappa2"], g["sf"], ego=ego)
    assert _close(_np(r["px"]), g["gx"]) and _close(_np(r["py"]), g["gy"]) and _close(_np(r["pyaw"]), g["gt"])
    assert _close(_np(r["pcos"]), np.cos(g["gt"])) and _close(_np(r["psin"]), np.sin(g["gt"]))
    assert _close(_np(r["end_xy"]), np.stack([g["gx"][:, -1], g["gy"][:, -1]]))
    # ego frame (no transform): the raw sample_spiral output; heading j is the sample before point j
    r0 = engine.sample_lattice(g["kappa1"], g["kappa2"], g["sf"], ego=None, want_trig=False, want_end=False)
    assert _close(_np(r0["px"]), g["x"]) and _close(_np(r0["py"]), g["y"]) and _close(_np(r0["pyaw"]), g["t"][:, :49])
    # one pose for the whole lattice (the planner's case: 7 goal states from one ego state)
    e0 = g["ego"][5]
    rb = engine.sample_lattice(g["kappa1"], g["kappa2"], g["sf"], ego=tuple(e0), want_trig=False)
    gx, gy, gt = wl.transform_to_global(g["x"], g["y"], g["t"], np.full(96, e0[0]), np.full(96, e0[1]), np.full(96, e0[2]))
    assert _close(_np(rb["px"]), gx) and _close(_np(rb["py"]), gy) and _close(_np(rb["pyaw"]), gt)


def test_device_resident_lattice_pipeline_cfg3(engine):
    """Config 3 generated on the device from its 4,096 spiral parameter triples, checked and scored without the
    paths ever visiting the host: flags and chosen index equal the host pipeline's on this batch."""
    w = wl.config3_lattice()
    par = w["spiral_params"]
    r = engine.sample_lattice(par[0], par[1], par[2], ego=w["ego"])
    assert _close(_np(r["px"]), w["px"], 1e-11) and _close(_np(r["py"]), w["py"], 1e-11) and _close(_np(r["pyaw"]), w["pyaw"])
    # the device-generated paths checked from their device-generated yaws: bit-exact against the oracle on the SAME
    # paths (whose circle centres carry numpy's cos / sin of those yaws)
    free = engine.collision_check_batch(r["px"], r["py"], r["pyaw"], w["obstacles"], OFF, RAD)
    ref_dev, _, _ = c_oracle.collision_check(_np(r["px"]), _np(r["py"]), _np(r["pyaw"]), w["obstacles"], OFF, RAD)
    assert np.array_equal(_np(free).astype(bool), ref_dev)
    # against the HOST-generated paths (coordinates differ by ~1e-12): a flag may differ only where an obstacle point
    # sits within 1e-9 m of a circle -- stated on the clearance, not as an allowance on the count
    ref, clr, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True)
    diff = _np(free).astype(bool) != ref
    print(f"device-generated lattice: {int(diff.sum())}/{len(ref)} flags differ from the host-generated pipeline")
    assert np.all(np.abs(clr[diff]) < 1e-9)
    mode = host_norm2_mode()
    best = engine.select_best_path_index_batch(r["end_xy"][0], r["end_xy"][1], free, w["goal"], W, norm_mode=mode)
    want, _ = c_oracle.select_best(_np(r["end_xy"])[0], _np(r["end_xy"])[1], _np(free), w["goal"], W, mode)
    assert best == want
    if not diff.any():
        want_host, _ = c_oracle.select_best(w["px"][:, -1], w["py"][:, -1], ref, w["goal"], W, mode)
        print(f"best index device pipeline {best}, host pipeline {want_host}")


def test_lattice_edge_cases(engine):
    z = engine.sample_lattice(np.zeros(0), np.zeros(0), np.zeros(0), ego=(0.0, 0.0, 0.0))
    assert z["px"].shape == (0, 49)
    r = engine.sample_lattice([0.0], [0.0], [30.0], ego=(1.0, 2.0, 0.0), n_samples=4)      # straight line, 3 points
    assert np.allclose(_np(r["px"])[0], [11.0, 21.0, 31.0]) and np.allclose(_np(r["py"])[0], 2.0)
    big = engine.sample_lattice(np.linspace(-0.05, 0.05, 1000), np.zeros(1000), np.full(1000, 25.0), n_samples=200)
    assert big["px"].shape == (1000, 199) and np.isfinite(_np(big["px"])).all()
    with pytest.raises(ValueError):
        engine.sample_lattice([0.0], [0.0, 1.0], [30.0])


def test_spiral_optimisation_vs_scipy_and_oracle(engine, golden):
    """K7 against the literal reference (scipy L-BFGS-B results for 192 goal states) and the NumPy restatement."""
    from oracle import spiral_opt_numpy as so
    g = golden("spiral_opt.npz")
    goals = g["goals"]
    r = engine.optimize_spirals(goals[:, 0], goals[:, 1], goals[:, 2])
    p = _np(r["p"]).T
    f = _np(r["objective"])
    # same basin, at least as deep as where L-BFGS-B stopped; parameters to the accuracy it stops at
    assert np.all(f <= g["res_f"] + 1e-9 * np.maximum(1.0, g["res_f"]))
    scale = np.stack([np.ones(len(p)), np.ones(len(p)), p[:, 2]], 1)
    dp = np.abs(p - g["res_p"]) / scale
    print(f"K7 vs scipy L-BFGS-B: worst parameter difference {dp.max():.2e}; largest objective gain {(g['res_f'] - f).max():.2e}; "
          f"iterations max {int(_np(r['iterations']).max())}")
    assert dp.max() < 5e-4
    assert np.array_equal(_np(r["valid"]).astype(bool), g["valid"])
    # the objective the kernel reports is the reference's objective at its parameters
    assert np.abs(f - so.objective(p, goals)).max() < 1e-11
    # tight agreement with the restated solver (same iteration, numpy vs CUDA sincos)
    po = np.array([so.optimize(goal)[0] for goal in goals])
    assert (np.abs(p - po) / scale).max() < 1e-8
    # sampled end states of the device spirals hit the goals like the reference's do
    lat = engine.sample_lattice(r["p"][0], r["p"][1], r["p"][2], ego=None, want_trig=False)
    end = np.stack([_np(lat["end_xy"])[0], _np(lat["end_xy"])[1]], 1)
    assert np.abs(end - g["res_end"][:, :2]).max() < 1e-4


def test_plan_lattice_pipeline_matches_host_pipeline(engine):
    """Goal states -> optimised spirals -> paths -> flags -> best index, all on the device, against the same pipeline
    assembled from the oracles on the host."""
    from oracle import spiral_opt_numpy as so
    rng = np.random.default_rng(3)
    P = 448                                   # 64 ego poses' worth of 7-path lattices, planned as one batch from one pose
    gt = rng.uniform(-0.3, 0.3, P)
    base_x, base_y = rng.uniform(22.0, 38.0, P), rng.uniform(-2.0, 2.0, P)
    off = (np.arange(P) % 7 - 3) * 2.0
    goals = np.stack([base_x + off * np.cos(gt + np.pi / 2), base_y + off * np.sin(gt + np.pi / 2), gt])
    ego = (12.0, -7.0, 0.4)
    w = wl.config3_lattice(P=8, M=3000)
    obstacles = w["obstacles"] * 0.6 - 10.0
    best, out = engine.plan_lattice(goals, ego, obstacles, OFF, RAD, (45.0, 10.0), W)
    # host pipeline
    ph = np.array([so.optimize(goals[:, i])[0] for i in range(P)])
    x, y, t = wl.sample_spirals(ph[:, 0], ph[:, 1], ph[:, 2])
    gx, gy, gyaw = wl.transform_to_global(x, y, t, np.full(P, ego[0]), np.full(P, ego[1]), np.full(P, ego[2]))
    assert np.abs(_np(out["px"]) - gx).max() < 1e-6 and np.abs(_np(out["py"]) - gy).max() < 1e-6
    # flags: bit-exact against the oracle on the device's own paths; against the host pipeline's paths (which differ by
    # the two optimisers' ~1e-6) a flag may differ only where the clearance is below that difference
    dev_free, _, _ = c_oracle.collision_check(_np(out["px"]), _np(out["py"]), _np(out["pyaw"]), obstacles, OFF, RAD)
    assert np.array_equal(_np(out["free"]).astype(bool), dev_free & _np(out["valid"]).astype(bool))
    ref_free, clr, _ = c_oracle.collision_check(gx, gy, gyaw, obstacles, OFF, RAD, want_clearance=True)
    diff = _np(out["free"]).astype(bool) != ref_free
    print(f"plan_lattice: {int(ref_free.sum())}/{P} free, {int(diff.sum())} flags differ, best {best}")
    assert np.all(np.abs(clr[diff]) < 1e-5) and 0 < ref_free.sum() < P
    if not diff.any():
        want, _ = c_oracle.select_best(gx[:, -1], gy[:, -1], ref_free, (45.0, 10.0), W, host_norm2_mode())
        assert best == want and out["best_filtered"] == best          # every spiral of this batch is valid


def test_plan_lattice_drops_invalid_spirals_like_the_reference(engine, golden):
    """Goal sets with unreachable goals against the literal ``plan_paths`` -> ``transform_paths`` -> ``collision_check`` ->
    ``select_best_path_index`` (tests/golden/plan_invalid.npz): a dropped spiral is neither candidate nor penalty term, and
    the reference's index refers to the filtered list (local_planner.py:317-323, 367-379)."""
    g = golden("plan_invalid.npz")
    for c in range(int(g["n_cases"])):
        goals, ego, obs = g[f"c{c}_goals"], g[f"c{c}_ego"], g[f"c{c}_obstacles"]
        validity, flags, want = g[f"c{c}_validity"], g[f"c{c}_flags"], int(g[f"c{c}_best"])
        best, out = engine.plan_lattice(goals[:, :3].T.copy(), tuple(ego[:3]), obs, OFF, RAD, g[f"c{c}_goal_state"][:2],
                                        float(g["weight"]))
        valid = _np(out["valid"]).astype(bool)
        assert np.array_equal(valid, validity), c
        assert np.array_equal(_np(out["free"]).astype(bool)[valid], flags), c
        assert np.abs(_np(out["end_xy"]).T[valid] - g[f"c{c}_ends"]).max(initial=0.0) < 1e-4
        assert out["best_filtered"] == (None if want < 0 else want), (c, out["best_filtered"], want)
        if want >= 0:
            assert valid[best] and int(valid[:best].sum()) == want
        else:
            assert best is None
