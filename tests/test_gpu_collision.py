"""GPU parity of the lattice evaluation: collision booleans and selected indices must be BIT-EXACT.

Compared with the literal-reference golden vectors, with the bit-exact oracle on the full config-3
batch (4,096 paths x 49 points x 3 circles vs 10,000 obstacle points), and through the reference-named
classes / the pool seam.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import c_oracle, collision_numpy as cn
from python_motionplanning_b200 import CollisionChecker, ThreadPool, workloads as wl
from python_motionplanning_b200.host_numerics import host_norm2_mode

pytestmark = pytest.mark.gpu
OFF, RAD, W = list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT


def _np(t):
    return t.cpu().numpy()


@pytest.fixture(autouse=True, params=["auto", "screen", "fp64"])
def collision_mode(request, engine):
    """Every test runs through both arithmetic modes of the boolean kernel: the FP32-screened default and the
    all-FP64 kernel must give the same bits."""
    prev = engine.set_collision_mode(request.param)
    yield request.param
    engine.set_collision_mode(prev)


def test_collision_subsample_vs_literal(engine, golden):
    g = golden("collision_cfg3_sub.npz")
    free = engine.collision_check_batch(g["px"], g["py"], g["pyaw"], g["obstacles"], OFF, RAD)
    assert np.array_equal(_np(free).astype(bool), g["free"])


def test_collision_kat3_and_boundary_vs_literal(engine, golden):
    g = golden("collision_cfg3_sub.npz")
    cc = CollisionChecker(OFF, RAD, W, engine=engine)
    path = [[float(i) for i in range(1, 50)], [0.0] * 49, [0.0] * 50]          # 50 yaws: only 49 are read
    for obs, cnt, want in zip(g["kat3_obstacles"], g["kat3_counts"], g["kat3_free"]):
        got = cc.collision_check(path, obs[:cnt].tolist())
        assert isinstance(got, bool) and got == bool(want)
    # obstacle points a few ulps either side of the radius (3,000 literal-reference verdicts)
    n = len(g["bnd_x"])
    got = np.zeros(n, dtype=bool)
    for i in range(n):
        f = engine.collision_check_batch(g["bnd_x"][i:i + 1, None], g["bnd_y"][i:i + 1, None], g["bnd_yaw"][i:i + 1, None],
                                         np.array([[g["bnd_ox"][i], g["bnd_oy"][i]]]), OFF, RAD)
        got[i] = bool(f.item())
    assert np.array_equal(got, g["bnd_free"])


def test_collision_cfg3_full_bit_exact(engine):
    """Full config 3 against the bit-exact C oracle, plus min-clearance and the early-exit-free mode."""
    w = wl.config3_lattice()
    free = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD)
    ref, _, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD)
    assert np.array_equal(_np(free).astype(bool), ref)
    assert 0.2 <= ref.mean() <= 0.8
    free2, clr = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True)
    ref2, clr_ref, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True)
    assert np.array_equal(_np(free2).astype(bool), ref) and np.array_equal(ref2, ref)
    assert np.array_equal(_np(clr), clr_ref)                     # same roundings -> identical doubles
    # the default call resolved only the candidate points (about one per path) with host trig; every yaw on the host gives
    # the same doubles
    assert 0 < engine.last_clearance_candidates <= 2 * len(ref)
    free_h2, clr_h = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True,
                                                  clearance_trig="host")
    assert np.array_equal(_np(free_h2).astype(bool), ref) and np.array_equal(_np(clr_h), clr_ref)
    # clearance with the yaws kept on the device: flags still the proven bit-exact ones, the clearance within a few ulp of a
    # circle centre (1e-12 m stated; the device's sincos moves a centre by ~2e-16 m)
    free3, clr3 = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True,
                                               clearance_trig="device")
    assert np.array_equal(_np(free3).astype(bool), ref)
    d = np.abs(_np(clr3) - clr_ref)
    print(f"device-trig clearance: max |diff| {d.max():.2e} m, {np.count_nonzero(d)} of {d.size} paths differ")
    assert d.max() < 1e-12
    # the default call above sent the YAWS to the device (proven verdicts + host-resolved leftovers); the former default
    # -- numpy cos / sin of every yaw on the host -- and caller-supplied trig give the same flags, bit for bit
    und = engine.last_collision_undecided
    free_h = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, host_trig=True)
    trig = engine.path_trig(w["pyaw"], w["px"].shape[1])
    free_t = engine.collision_check_batch(w["px"], w["py"], None, w["obstacles"], OFF, RAD, trig=trig)
    assert np.array_equal(_np(free_h).astype(bool), ref) and np.array_equal(_np(free_t).astype(bool), ref)
    # unproven device-only trigonometry (A/B): report
    free3 = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, device_trig=True)
    mism = int((_np(free3).astype(bool) != ref).sum())
    print(f"config 3: free fraction {ref.mean():.3f}; path points resolved on the host {und}; device-only-trig mismatches {mism}/{len(ref)}")
    assert und <= 64            # an obstacle point within ~1e-13 m of a circle is rare


def test_collision_yaw_mode_boundary_bands_resolved_on_host(engine):
    """Obstacle points placed a few FP64 ulps to 1e-12 either side of a circle computed with numpy's cos / sin: the
    device cannot prove those verdicts, must list the path points (never guess) and the host-resolved flags must equal
    the oracle's -- in all three arithmetic modes, for per-call ``mode=`` and with yaws up to 1e4 rad."""
    rng = np.random.default_rng(17)
    P, n = 96, 49
    for scale, yaw_scale in ((1.0, np.pi), (1.0e3, 50.0), (1.0e6, 1.0e4)):
        px, py = rng.uniform(-100, 100, (P, n)) * scale / 1.0, rng.uniform(-100, 100, (P, n)) * scale / 1.0
        pyaw = rng.uniform(-yaw_scale, yaw_scale, (P, n + 1))
        k = rng.integers(0, 3, P)
        j = rng.integers(0, n, P)
        ang = rng.uniform(-np.pi, np.pi, P)
        off = np.array(OFF)[k]
        c, s_ = np.cos(pyaw[np.arange(P), j]), np.sin(pyaw[np.arange(P), j])
        cx, cy = px[np.arange(P), j] + off * c, py[np.arange(P), j] + off * s_
        eps = np.where(np.arange(P) % 2 == 0, 1, -1) * 10.0 ** rng.uniform(-16, -11.5, P)
        rad = np.array(RAD)[k] * (1.0 + eps)
        near = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)       # one near-boundary point per path
        far = np.stack([rng.uniform(-100, 100, 5000) * scale + 1.0e3 * scale + 1.0e3, rng.uniform(-100, 100, 5000) * scale], 1)
        obs = np.concatenate([near, far])
        ref, _, _ = c_oracle.collision_check(px, py, pyaw, obs, OFF, RAD)
        seen = 0
        for mode in ("auto", "screen", "fp64"):
            got = _np(engine.collision_check_batch(px, py, pyaw, obs, OFF, RAD, mode=mode)).astype(bool)
            seen = max(seen, engine.last_collision_undecided)
            assert np.array_equal(got, ref), (scale, mode, int((got != ref).sum()))
        assert seen > 0, "the band was never exercised"
        print(f"scale {scale:g}: {int(ref.sum())}/{P} free, up to {seen} path points resolved on the host")
    # more undecided points than the list holds -> the call falls back to host trig for every yaw, same flags
    cap = type(engine)._UNDECIDED_CAPACITY
    try:
        type(engine)._UNDECIDED_CAPACITY = 4
        engine._und = None
        got = _np(engine.collision_check_batch(px, py, pyaw, obs, OFF, RAD)).astype(bool)
        assert np.array_equal(got, ref) and engine.last_collision_undecided > 4
    finally:
        type(engine)._UNDECIDED_CAPACITY = cap
        engine._und = None


def test_collision_permutation_and_tiling_invariance(engine):
    """The verdict does not depend on obstacle order / tile boundaries (1,024-point tiles)."""
    w = wl.config3_lattice(P=512, M=3000)
    base = _np(engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD))
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(w["obstacles"]))
    assert np.array_equal(base, _np(engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"][perm], OFF, RAD)))
    for M in (1, 1023, 1024, 1025, 2049):
        ref, _, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"][:M], OFF, RAD)
        got = _np(engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"][:M], OFF, RAD)).astype(bool)
        assert np.array_equal(got, ref), M
    # other circle counts / radii
    for off, rad in (([0.0], [2.0]), ([-1.0, 0.5, 2.0, 3.5, 5.0], [1.0, 1.2, 0.8, 1.5, 0.3])):
        ref, _, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], off, rad)
        got = _np(engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], off, rad)).astype(bool)
        assert np.array_equal(got, ref)


def test_select_best_vs_literal_and_oracle(engine, golden):
    g = golden("collision_cfg3_sub.npz")
    w = wl.config3_lattice()
    ex, ey = w["px"][:, -1].copy(), w["py"][:, -1].copy()
    mode = host_norm2_mode()
    free512 = g["sel_free512"]
    if mode == int(g["norm2_mode"]):
        for lo, n, best in g["sel_cases"]:      # literal reference results (same BLAS closed form as this host)
            got = engine.select_best_path_index_batch(ex[lo:lo + n], ey[lo:lo + n], free512[lo:lo + n], g["goal"], W)
            assert got == (None if best < 0 else int(best))
    # full P = 4,096 against the C oracle in every rounding mode, scores bit for bit
    free, _, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD)
    for m in (0, 1, 2):
        want, scores = c_oracle.select_best(ex, ey, free, w["goal"], W, m)
        got, sc = engine.select_best_path_index_batch(ex, ey, free, w["goal"], W, norm_mode=m, want_scores=True)
        assert got == want and np.array_equal(_np(sc), scores)
    # host-mode result equals the literal-order NumPy oracle (which calls np.linalg.norm itself) on 512 paths
    want = cn.select_best_path_index(ex[:512], ey[:512], free[:512], w["goal"], W)
    assert engine.select_best_path_index_batch(ex[:512], ey[:512], free[:512], w["goal"], W) == want


def test_select_best_kat4_ties(engine, golden):
    g = golden("collision_cfg3_sub.npz")
    cc = CollisionChecker(OFF, RAD, W, engine=engine)
    paths = [[[float(i) for i in range(1, 50)], [yv] * 49, [0.0] * 49] for yv in (-4.0, -2.0, 0.0, 2.0, 4.0)]
    for flags, best in zip(g["kat4_flags"], g["kat4_best"]):
        got = cc.select_best_path_index(paths, [bool(f) for f in flags], [49, 0, 25])
        assert got == (None if best < 0 else int(best))
    assert cc.select_best_path_index([], [], [0, 0, 0]) is None


def test_planner_seam_closed_loop_lattices(engine, golden):
    """Config 1: the reference's own planner lattices through the pool seam, flags + index per frame."""
    g = golden("closedloop_cfg1.npz")
    cc = CollisionChecker(OFF, RAD, W, engine=engine)
    obstacle = g["obstacle_xy"].tolist()
    for k, f in enumerate(g["plan_path_frames"]):
        paths = [[g["plan_paths"][k, i, 0].tolist(), g["plan_paths"][k, i, 1].tolist(), g["plan_paths"][k, i, 2].tolist()]
                 for i in range(7)]
        # exactly the call local_planner.py:370-372 makes
        pool = ThreadPool(processes=len(paths))
        flags = pool.starmap(cc.collision_check, zip(paths, itertools.repeat(obstacle)))
        assert flags == [bool(x) for x in g["plan_flags"][f]], f
        goal = list(g["plan_goal"][f]) + [25.0]
        best = cc.select_best_path_index(paths, flags, goal)
        assert (-1 if best is None else best) == int(g["plan_best"][f]), f
    with pytest.raises(ValueError):
        ThreadPool(processes=0)               # keeps the planner's `except ValueError -> [True]*7` path


def test_collision_edge_cases(engine):
    cc = CollisionChecker(OFF, RAD, W, engine=engine)
    path = [[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]
    assert cc.collision_check(path, []) is True                       # empty obstacle list
    assert cc.collision_check([[], [], []], [[1.0, 0.0]]) is True     # empty path
    assert cc.collision_check_paths([], [[1.0, 0.0]]) == []
    # ragged lengths in one batch
    long = [[float(i) for i in range(10)], [0.0] * 10, [0.0] * 10]
    flags = cc.collision_check_paths([path, long, [[], [], []]], [[8.0, 0.5]])
    assert flags == [True, False, True]
    # dist == r is free, one ulp inside is a collision
    assert cc.collision_check([[0.0], [0.0], [0.0]], [[4.5, 0.0]]) is True
    assert cc.collision_check([[0.0], [0.0], [0.0]], [[np.nextafter(4.5, 0.0), 0.0]]) is False
    # zero-size launches through the engine
    z = np.zeros((0, 49))
    assert engine.collision_check_batch(z, z, z, np.zeros((5, 2)), OFF, RAD).numel() == 0
    f = engine.collision_check_batch(np.zeros((3, 4)), np.zeros((3, 4)), np.zeros((3, 4)), np.zeros((0, 2)), OFF, RAD)
    assert _np(f).tolist() == [1, 1, 1]
    with pytest.raises(ValueError):
        engine.collision_check_batch(np.zeros((3, 4)), np.zeros((3, 4)), np.zeros((3, 4)), np.zeros((2, 2)), [0.0] * 9, [1.0] * 9)


def _ring(rng, cx, cy, r, n, rel):
    """n points at distance r*(1+rel_i) from (cx, cy) in random directions."""
    th = rng.uniform(-np.pi, np.pi, n)
    d = r * (1.0 + rel)
    return np.stack([cx + d * np.cos(th), cy + d * np.sin(th)], axis=1)


@pytest.mark.parametrize("scale_shift", [(1.0, 0.0), (1.0, 1.0e3), (1.0, 1.0e6), (1.0, 1.0e9), (1.0e-3, 0.0), (1.0e4, -3.0e7)])
def test_collision_filter_adversarial_bands(engine, scale_shift):
    """Obstacle points packed around the circle boundaries at every scale the FP32 screen has to get right:
    a few FP64 ulps, around the FP32 rounding band (1e-8 .. 1e-4 relative), and clearly inside / outside."""
    scale, shift = scale_shift
    rng = np.random.default_rng(7)
    w = wl.config3_lattice(P=256, M=64)
    px, py, pyaw = w["px"] * scale + shift, w["py"] * scale + shift, w["pyaw"]
    off = [o * scale for o in OFF]
    rad = [r * scale for r in RAD]
    P, n = px.shape
    obs = []
    rels = np.concatenate([np.arange(-8, 9) * 2.0 ** -52, [-1e-4, -1e-5, -1e-6, -1e-7, -1e-8, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4]])
    for p in range(0, P, 2):                       # every second path gets boundary points; the others stay far from them
        j = int(rng.integers(0, n))
        k = int(rng.integers(0, 3))
        c, s_ = np.cos(pyaw[p, j]), np.sin(pyaw[p, j])
        cx, cy = px[p, j] + off[k] * c, py[p, j] + off[k] * s_
        obs.append(_ring(rng, cx, cy, rad[k], len(rels), rng.permutation(rels)))
    obs = np.concatenate(obs)
    ref, _, _ = c_oracle.collision_check(px, py, pyaw, obs, off, rad)
    got = _np(engine.collision_check_batch(px, py, pyaw, obs, off, rad)).astype(bool)
    assert np.array_equal(got, ref)
    assert 0 < ref.sum() < P                       # the case exercises both verdicts
    # exact minimum clearance (FP32 screen for the candidate set + exact FP64 on the candidates): the same doubles
    ref2, clr_ref, _ = c_oracle.collision_check(px, py, pyaw, obs, off, rad, want_clearance=True)
    free2, clr = engine.collision_check_batch(px, py, pyaw, obs, off, rad, want_clearance=True)
    assert np.array_equal(_np(free2).astype(bool), ref2) and np.array_equal(_np(clr), clr_ref)
    # one path at a time as well (different tile / origin per call)
    for p in range(0, 32):
        ref1, _, _ = c_oracle.collision_check(px[p:p + 1], py[p:p + 1], pyaw[p:p + 1], obs, off, rad)
        got1 = _np(engine.collision_check_batch(px[p:p + 1], py[p:p + 1], pyaw[p:p + 1], obs, off, rad)).astype(bool)
        assert np.array_equal(got1, ref1), p


@pytest.mark.parametrize("size", [(128, 512), (192, 1024)], ids=["single-launch", "broad-phase"])
def test_collision_filter_exceptional_values(engine, size):
    """NaN / Inf / huge coordinates and degenerate radii: the screened kernels must fall back to the exact
    sequence wherever single precision cannot bound its own error.  The second size is past the 4 M point-pair
    threshold, i.e. flags AND min-clearance take the bounding-box broad phase (collision_cull_kernel /
    clearance_cull_kernel)."""
    rng = np.random.default_rng(11)
    w = wl.config3_lattice(P=size[0], M=size[1])
    px, py, pyaw, obs = w["px"].copy(), w["py"].copy(), w["pyaw"], w["obstacles"].copy()

    def both(px, py, obs, off=OFF, rad=RAD, what=""):
        ref, _, _ = c_oracle.collision_check(px, py, pyaw, obs, off, rad)
        got = _np(engine.collision_check_batch(px, py, pyaw, obs, off, rad)).astype(bool)
        assert np.array_equal(got, ref), f"{what}: {int((got != ref).sum())} flags differ"
        # min-clearance: the screened kernel returns the doubles of the all-FP64 kernel, bit for bit, whatever the input
        f1, c1 = engine.collision_check_batch(px, py, pyaw, obs, off, rad, want_clearance=True)
        prev = engine.set_collision_mode("fp64")
        f0, c0 = engine.collision_check_batch(px, py, pyaw, obs, off, rad, want_clearance=True)
        engine.set_collision_mode(prev)
        assert torch.equal(f1, f0) and torch.equal(c1.view(torch.int64), c0.view(torch.int64)), f"{what}: clearance differs"
        return ref

    base = both(px, py, obs)
    assert 0 < base.sum() < len(base)
    for bad in (np.nan, np.inf, -np.inf, 1.0e300, -1.0e39, 3.5e38):
        o = obs.copy()
        o[0, 0] = bad                                         # the shift origin itself is unusable
        both(px, py, o, what=f"origin {bad}")
        o = obs.copy()
        o[rng.integers(1, len(o), 20), rng.integers(0, 2, 20)] = bad
        both(px, py, o, what=f"obstacles {bad}")
        x = px.copy()
        x[rng.integers(0, size[0], 10), rng.integers(0, 49, 10)] = bad
        both(x, py, obs, what=f"path {bad}")
    # both far away and close together: differences are small, magnitudes beyond FP32 range
    both(px + 1.0e39, py, obs + np.array([1.0e39, 0.0]), what="1e39")
    both(px + 1.0e15, py - 1.0e15, obs + np.array([1.0e15, -1.0e15]), what="1e15")
    # degenerate radii / offsets
    for rad in ([0.0, 0.0, 0.0], [-1.0, 1.5, 0.0], [1.0e-9, 1.0e-12, 1.0e-300], [1.0e6, 1.5, 1.0e-3], [np.inf, 1.5, 1.5], [np.nan, 1.5, 1.5]):
        both(px, py, obs, OFF, rad, what=f"radii {rad}")
    # an obstacle exactly on a circle centre and on a path point
    o = obs.copy()
    o[5] = (px[3, 7], py[3, 7])
    both(px, py, o)


def test_calls_on_two_streams_do_not_share_scratch(engine):
    """the ABI is asynchronous on the caller's stream; the library's scratch areas are per (device, stream), so
    interleaved calls on two streams (broad-phase collision + min-clearance + select_best on each) give the results of the
    same calls made one after the other."""
    wa, wb = wl.config3_lattice(P=1024, M=6000), wl.config3_lattice(P=768, M=9000, seed=wl.SEED + 1)
    serial = []
    for w in (wa, wb):
        f = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD)
        f2, c = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, want_clearance=True)
        serial.append((f.clone(), c.clone()))
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    dev = {k: {n: engine.dev(w[n]) for n in ("px", "py", "obstacles")} for k, w in (("a", wa), ("b", wb))}
    trig = {"a": engine.path_trig(wa["pyaw"], 49), "b": engine.path_trig(wb["pyaw"], 49)}
    torch.cuda.synchronize()
    out = {"a": [], "b": []}
    for _ in range(6):                                   # interleave: neither stream waits for the other
        for k, st in (("a", sa), ("b", sb)):
            with torch.cuda.stream(st):
                x = dev[k]
                f, c = engine.collision_check_batch(x["px"], x["py"], None, x["obstacles"], OFF, RAD, trig=trig[k], want_clearance=True)
                f1 = engine.collision_check_batch(x["px"], x["py"], None, x["obstacles"], OFF, RAD, trig=trig[k])
                out[k].append((f1, c))
    torch.cuda.synchronize()
    for k, ref in (("a", serial[0]), ("b", serial[1])):
        for f1, c in out[k]:
            assert torch.equal(f1, ref[0]) and torch.equal(c.view(torch.int64), ref[1].view(torch.int64)), k


def test_broad_phase_spatial_sort_edge_cases(engine):
    """The broad phase sorts the obstacle points into Morton order of a grid over their bounding box: degenerate extents
    (all points identical, all on one line), fewer points than a chunk, shuffled order, and a list beyond the sorter's
    limit (which keeps the caller's order) all give the oracle's flags and min-clearance."""
    w = wl.config3_lattice(P=192, M=3000)
    px, py, pyaw = w["px"], w["py"], w["pyaw"]
    rng = np.random.default_rng(23)
    cases = {
        "identical": np.repeat(w["obstacles"][:1], 2000, axis=0),
        "vertical line": np.stack([np.full(2500, w["obstacles"][7, 0]), np.linspace(-20.0, 120.0, 2500)], 1),
        "horizontal line": np.stack([np.linspace(-20.0, 120.0, 2500), np.full(2500, w["obstacles"][9, 1])], 1),
        "shuffled": w["obstacles"][rng.permutation(3000)],
        "two far clusters": np.concatenate([w["obstacles"][:1500], w["obstacles"][1500:] + 1.0e6]),
    }
    for name, obs in cases.items():
        ref, clr_ref, _ = c_oracle.collision_check(px, py, pyaw, obs, OFF, RAD, want_clearance=True)
        got = _np(engine.collision_check_batch(px, py, pyaw, obs, OFF, RAD)).astype(bool)
        assert np.array_equal(got, ref), name
        f2, clr = engine.collision_check_batch(px, py, pyaw, obs, OFF, RAD, want_clearance=True)
        assert np.array_equal(_np(f2).astype(bool), ref) and np.array_equal(_np(clr), clr_ref), name
    # beyond kSortMaxPoints (131,072): the broad phase runs on the caller's order
    big = wl.config3_lattice(P=48, M=140000)
    ref, clr_ref, _ = c_oracle.collision_check(big["px"], big["py"], big["pyaw"], big["obstacles"], OFF, RAD, want_clearance=True)
    got = _np(engine.collision_check_batch(big["px"], big["py"], big["pyaw"], big["obstacles"], OFF, RAD)).astype(bool)
    f2, clr = engine.collision_check_batch(big["px"], big["py"], big["pyaw"], big["obstacles"], OFF, RAD, want_clearance=True)
    assert np.array_equal(got, ref) and np.array_equal(_np(f2).astype(bool), ref) and np.array_equal(_np(clr), clr_ref)
