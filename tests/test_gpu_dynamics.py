"""GPU parity of the 7-DoF RK4 path: CUDA kernels (through the C ABI) vs the oracle and the golden vectors.

Contract (BASELINE.json north_star): FP64 within 1e-9 relative on every state component after N steps,
evaluated as |gpu - ref| <= 1e-9 * max(|ref|, 1); FP32 within a measured drift bound.
"""
import numpy as np
import pytest
import torch

from conftest import REL_TOL_F64, rel_err
from oracle import c_oracle, mpc_numpy, planar_numpy as pn
from python_motionplanning_b200 import VehicleModel, VehicleParameters, workloads as wl

pytestmark = pytest.mark.gpu
DT = 1e-4


@pytest.fixture(autouse=True, params=["auto", "closed_form"])
def friction_mode(request, engine):
    """Every test runs with the tabulated friction path of the fast kernels and with the closed form."""
    prev = engine.set_friction_mode(request.param)
    yield request.param
    engine.set_friction_mode(prev)


def _params(D=1.0):
    p = VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = D
    return p


def _c_params(D=1.0):
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (D,) * 4
    return par


def test_planar_model_vs_literal_golden(engine, golden):
    g = golden("planar_model.npz")
    engine.set_params(_params())
    sd, misc, out = engine.planar_model_batch(g["states"].T, g["torque"].T, g["mu_max"].T, g["delta"].T,
                                              g["ax_prev"], g["ay_prev"])
    assert rel_err(sd.cpu().numpy().T, g["state_dot"]).max() < REL_TOL_F64
    assert rel_err(misc.cpu().numpy().T, g["misc"]).max() < REL_TOL_F64
    assert rel_err(out.cpu().numpy().T, g["outputs"]).max() < REL_TOL_F64
    # zero-slip state: exactly zero forces (the s == 0 branch, vehicle_model.py:309-348)
    assert list(sd.cpu().numpy()[:, 1]) == [0, 0, 0, 0, 0, 0, 0, 0, 25, 0]


def test_rk4_single_step_vs_literal_golden(engine, golden):
    g = golden("planar_model.npz")
    engine.set_params(_params())
    n = len(g["states"])
    s0 = np.concatenate([g["states"].T, g["ax_prev"][None], g["ay_prev"][None]])
    res = engine.rollout(s0, g["delta"].T[None], g["torque"].T[None], float(g["dt"]), 1, mu=g["mu_max"].T,
                         store_stride=1, want_aux=True)
    end = res.state_end.cpu().numpy()
    aux = res.aux.cpu().numpy()[0]
    assert rel_err(end[:10].T, g["rk4_state"]).max() < REL_TOL_F64
    assert rel_err(end[10:].T, g["rk4_axay"]).max() < REL_TOL_F64
    assert rel_err(aux[:10].T, g["rk4_state_dot"]).max() < REL_TOL_F64
    assert rel_err(aux[10:].T, g["rk4_outputs"]).max() < REL_TOL_F64
    assert np.array_equal(res.traj.cpu().numpy()[0], end[:10]) and n == end.shape[1]


def test_rollout_cfg2_subsample_vs_literal(engine, golden):
    """256 rollouts x 500 steps of config 2 against the literal reference."""
    g = golden("rollout_cfg2_sub.npz")
    engine.set_params(_params())
    res = engine.rollout(g["state0"], g["delta"], g["torque"], float(g["dt"]), 500, hold=int(g["hold"]), store_stride=1)
    traj = res.traj.cpu().numpy()
    worst = 0.0
    for k, n in enumerate(g["check_steps"]):
        e = rel_err(traj[n - 1], g["states"][:, k, :10].T).max()
        worst = max(worst, e)
        assert e < REL_TOL_F64, (n, e)
    assert rel_err(res.state_end.cpu().numpy()[10:], g["states"][:, -1, 10:].T).max() < REL_TOL_F64
    print(f"config-2 subsample vs literal reference: worst rel err {worst:.3e}")


def test_rollout_cfg2_full_vs_c_oracle(engine):
    """Full config 2 (65,536 x 500, FP64) against the C oracle on identical inputs."""
    B, N = 65536, 500
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    engine.set_params(_params())
    res = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=50)
    ref = c_oracle.rollout(s0, d, t, _c_params(), DT, N, hold=wl.HOLD, store_stride=50)
    e = rel_err(res.traj.cpu().numpy(), ref["traj"])
    assert e.max() < REL_TOL_F64, e.max()
    assert rel_err(res.state_end.cpu().numpy(), ref["state_end"]).max() < REL_TOL_F64
    print(f"config 2 full: worst rel err {e.max():.3e}, per component {e.max(axis=(0, 2))}")


def test_closed_loop_replay_cfg1(engine, golden):
    """Config 1: replay the reference's recorded closed-loop controls (ZOH-10), all 40,000 steps."""
    g = golden("closedloop_cfg1.npz")
    engine.set_params(_params())
    n = len(g["delta"]) * 10
    res = engine.rollout(g["state0"][:, None], g["delta"][:, None, None], g["torque"][:, None, None], float(g["dt"]),
                         n, hold=10, store_stride=10)
    traj = res.traj.cpu().numpy()[:, :, 0]
    e = rel_err(traj, g["state_every10"])
    assert e.max() < REL_TOL_F64, e.max()
    assert rel_err(res.state_end.cpu().numpy()[10:, 0], g["axay_every10"][-1]).max() < 1e-8
    print(f"closed-loop replay 40,000 steps: worst rel err {e.max():.3e}")


def test_scalar_vehicle_model_surface(golden):
    """The reference-named scalar calls: return structure, p.D mutation, and values (first frame)."""
    g = golden("closedloop_cfg1.npz")
    vm = VehicleModel(2.906, np.deg2rad(30), float(g["dt"]))
    p = VehicleParameters()
    state = list(g["state0"][:10])
    ax = ay = 0
    for i in range(100):
        d, t = g["delta"][i // 10], g["torque"][i // 10]
        r = vm.planar_model_RK4(state, [t, t, t, t], [1.0, 1.0, 1.0, 1.0], [d, d, 0, 0], p, ax, ay)
        assert len(r) == 9 and r[0].shape == (10,) and r[5].shape == (10,) and r[6].shape == (18,)
        state, ax, ay = r[0], r[7], r[8]
        assert r[1] == state[8] and r[2] == state[9] and r[3] == state[7] and r[4] == state[0]
        assert rel_err(state, g["first_frame_states"][i]).max() < REL_TOL_F64
        assert rel_err(r[5], g["first_frame_sdot"][i]).max() < REL_TOL_F64
        assert rel_err(r[6], g["first_frame_outputs"][i]).max() < REL_TOL_F64
    assert (p.DFL, p.DFR, p.DRL, p.DRR) == (1.0, 1.0, 1.0, 1.0)
    k = golden("planar_model.npz")
    r = vm.planar_model(list(k["states"][0]), list(k["torque"][0]), list(k["mu_max"][0]), list(k["delta"][0]), p,
                        0.4, -0.7)
    assert len(r) == 8 and r[0].shape == (10,) and r[5].shape == (18,)
    assert rel_err(r[0], k["state_dot"][0]).max() < REL_TOL_F64
    assert rel_err([r[1], r[2], r[3], r[4], r[6], r[7]], k["misc"][0]).max() < REL_TOL_F64
    assert (p.DFL, p.DFR, p.DRL, p.DRR) == tuple(k["mu_max"][0])


def test_rollout_is_resumable_bitwise(engine):
    """Two launches of 250 steps (step0 = 0, 250) equal one launch of 500 bit for bit."""
    s0, d, t = wl.config2_rollouts(B=4096, n_steps=500)
    engine.set_params(_params())
    one = engine.rollout(s0, d, t, DT, 500, hold=wl.HOLD, store_stride=10)
    a = engine.rollout(s0, d, t, DT, 250, hold=wl.HOLD, store_stride=10)
    b = engine.rollout(a.state_end, d, t, DT, 250, hold=wl.HOLD, store_stride=10, step0=250)
    assert torch.equal(one.state_end, b.state_end)
    assert torch.equal(one.traj, torch.cat([a.traj, b.traj]))


def test_rollout_layout_variants_agree(engine):
    """4-channel controls, per-rollout mu_max and param_set reproduce the front-steer fast path."""
    B, N = 2048, 200
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    engine.set_params(_params())
    fast = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=0)
    z = np.zeros_like(d)
    d4 = np.concatenate([d, d, z, z], axis=1)
    t4 = np.repeat(t, 4, axis=1)
    gen = engine.rollout(s0, d4, t4, DT, N, hold=wl.HOLD)
    mu = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, mu=np.ones((4, B)))
    ps = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, param_set=np.zeros(B, dtype=np.int32))
    for other in (gen, mu, ps):
        assert rel_err(other.state_end.cpu().numpy(), fast.state_end.cpu().numpy()).max() < 1e-12
    # broadcast controls: one sequence for every rollout
    bc = engine.rollout(s0, d[:, :, :1], t[:, :, :1], DT, N, hold=wl.HOLD, ctrl_broadcast=True)
    rep = engine.rollout(s0, np.repeat(d[:, :, :1], B, 2), np.repeat(t[:, :, :1], B, 2), DT, N, hold=wl.HOLD)
    assert torch.equal(bc.state_end, rep.state_end)


def test_param_sweep_cfg5_small_f64_and_f32_drift(engine):
    """Config 5 (reduced: 16 tyre sets x 256 manoeuvres x 500 steps): FP64 parity + measured FP32 drift."""
    n_sets, n_man, N = 16, 256, 500
    sets, s0, d, t, pset = wl.config5_sweep(n_sets=n_sets, n_man=n_man)
    p = VehicleParameters()
    for w in ("FL", "FR", "RL", "RR"):
        setattr(p, "B" + w, sets[:, 0])
        setattr(p, "C" + w, sets[:, 1])
        setattr(p, "D" + w, sets[:, 2])
    assert engine.set_params(p) == n_sets
    r64 = engine.rollout(s0, d, t, DT, N, hold=N, param_set=pset, store_stride=1)
    op = pn.VehicleParams()
    for w in ("FL", "FR", "RL", "RR"):
        setattr(op, "B" + w, sets[:, 0])
        setattr(op, "C" + w, sets[:, 1])
        setattr(op, "D" + w, sets[:, 2])
    ref = c_oracle.rollout(s0, d, t, c_oracle.make_params(op), DT, N, hold=N, param_set=pset, store_stride=1)
    e = rel_err(r64.traj.cpu().numpy(), ref["traj"])
    assert e.max() < REL_TOL_F64, e.max()
    r32 = engine.rollout(s0, d, t, DT, N, hold=N, param_set=pset, store_stride=1, dtype="f32")
    t32 = r32.traj.cpu().numpy().astype(np.float64)
    drift = rel_err(t32, ref["traj"])
    report = {n: float(drift[n - 1].max()) for n in (1, 10, 100, 500)}
    print("FP32 drift (max rel, floor 1) at steps 1/10/100/500:", report)
    # stated, measured bound (profiles/r01_fp32_drift_cfg5.json: full config 5 measures 2.6e-5 max / 1.2e-5
    # p99 over the 10 states after 500 steps, 1.4e-7 after one): FP32 stays within 2e-4 of FP64 over 500 steps
    assert report[500] < 2e-4 and report[1] < 1e-6


def test_generic_tabulated_step_mixed_blocks_and_mu(engine):
    """Per-rollout parameter sets + per-wheel mu_max on the generic kernel: blocks whose 64 rollouts share one set take
    the tabulated step (D = 1 tables, mu_max scales the normal load), a set whose tyres differ between the axles has no
    table, and blocks that mix sets take the closed form -- all inside one launch, all against the C oracle."""
    n_sets, B, N = 6, 64 * 12 - 13, 300        # the last block is partial
    rng = np.random.default_rng(11)
    sets = np.stack([rng.uniform(8.0, 25.0, n_sets), rng.uniform(1.2, 1.9, n_sets), rng.uniform(0.3, 1.2, n_sets)], 1)
    p, op = VehicleParameters(), pn.VehicleParams()
    for w in ("FL", "FR", "RL", "RR"):
        for q in (p, op):
            setattr(q, "B" + w, sets[:, 0] * (1.1 if w[0] == "R" else 1.0) ** (np.arange(n_sets) == 3))   # set 3: axles differ
            setattr(q, "C" + w, sets[:, 1].copy())
            setattr(q, "D" + w, sets[:, 2].copy())
    assert engine.set_params(p) == n_sets
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    pset = np.concatenate([np.repeat(np.arange(6, dtype=np.int32), 64),            # six uniform blocks (one without table)
                           rng.integers(0, n_sets, B - 6 * 64).astype(np.int32)])  # six mixed blocks
    mu = rng.uniform(0.4, 1.1, (4, B))
    for kw in (dict(param_set=pset), dict(param_set=pset, mu=mu), dict(mu=mu)):
        got = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=10, **kw)
        ref = c_oracle.rollout(s0, d, t, c_oracle.make_params(op), DT, N, hold=wl.HOLD, store_stride=10, **kw)
        e = rel_err(got.traj.cpu().numpy(), ref["traj"])
        assert e.max() < REL_TOL_F64, (sorted(kw), e.max())


def test_mpc_controls_cost_argmin(engine):
    """Config 4 (reduced): Philox control sampling, running cost and device argmin vs restatements."""
    cfg = wl.config4_mpc(B=8192, n_steps=100)
    B, N = cfg["B"], cfg["n_steps"]
    engine.set_params(_params())
    d, t = engine.mpc_sample_controls(B, N, cfg["seed"])
    dn, tn = mpc_numpy.sample_controls(B, N, cfg["seed"])
    assert np.abs(d.cpu().numpy() - dn).max() < 1e-15 and np.abs(t.cpu().numpy() - tn).max() < 1e-11
    # sharded sampling (global rollout offset) reproduces the unsharded sequences bit for bit
    d2, _ = engine.mpc_sample_controls(B // 2, N, cfg["seed"], rollout0=B // 2)
    assert torch.equal(d2, d[:, :, B // 2:])
    s0 = np.repeat(cfg["state0"][:, None], B, axis=1)
    res = engine.rollout(s0, d, t, DT, N, hold=1, cost_ref=cfg["cost_ref"], w_u=cfg["w_u"], u_ref=cfg["u_ref"],
                         store_stride=1)
    ref = c_oracle.rollout(s0, d.cpu().numpy(), t.cpu().numpy(), _c_params(), DT, N, hold=1, store_stride=1,
                           cost_ref=cfg["cost_ref"], w_u=cfg["w_u"], u_ref=cfg["u_ref"])
    assert rel_err(res.traj.cpu().numpy(), ref["traj"]).max() < REL_TOL_F64
    cost = res.cost.cpu().numpy()
    assert rel_err(cost, ref["cost"], 1e-12).max() < 1e-8
    # the cost equals the documented definition evaluated on the GPU's own trajectory
    assert rel_err(cost, mpc_numpy.rollout_cost(res.traj.cpu().numpy(), cfg["cost_ref"], cfg["w_u"], cfg["u_ref"]),
                   1e-300).max() < 1e-13
    mn, ix = engine.argmin(res.cost, index_offset=1000)
    assert int(ix.item()) == mpc_numpy.argmin_lowest(cost) + 1000 and float(mn.item()) == cost.min()
    # ties -> lowest index; NaN -> +inf; nothing finite -> -1
    c = torch.tensor([3.0, float("nan"), 1.0, 1.0, float("inf")], dtype=torch.float64)
    assert int(engine.argmin(c)[1].item()) == 2
    assert int(engine.argmin(torch.full((5000,), float("inf"), dtype=torch.float64))[1].item()) == -1
    big = torch.rand(3_000_000, dtype=torch.float64, device="cuda")
    big[2_345_678] = -1.0
    big[2_999_999] = -1.0
    assert int(engine.argmin(big)[1].item()) == 2_345_678


def test_rollout_to_host_pipeline(engine):
    """The end-to-end path (pinned host buffers, time-chunked, overlapped D2H) equals one device launch."""
    B, N = 8192, 200
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    engine.set_params(_params())
    dev = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=1)
    hs, hd, ht = (torch.from_numpy(a).pin_memory() for a in (s0, d, t))
    out = torch.empty(N, 10, B, dtype=torch.float64).pin_memory()
    end = engine.rollout_to_host(hs, hd, ht, DT, N, wl.HOLD, out, chunk_steps=50)
    assert torch.equal(out, dev.traj.cpu()) and torch.equal(end, dev.state_end)
    # end states only: controls uploaded chunk by chunk behind the kernels, no trajectory readback
    end_host = torch.empty(12, B, dtype=torch.float64).pin_memory()
    for chunk in (50, 80, 200):
        end_host.zero_()
        engine.rollout_endstate_to_host(hs, hd, ht, DT, N, wl.HOLD, end_host, chunk_steps=chunk)
        assert torch.equal(end_host, dev.state_end.cpu())


def test_abi_error_behaviour(engine):
    s0, d, t = wl.config2_rollouts(B=64, n_steps=20)
    engine.set_params(_params())
    with pytest.raises(ValueError):
        engine.rollout(s0, d[:, :, :10], t, DT, 20, hold=wl.HOLD)          # controls do not cover the batch
    with pytest.raises(ValueError):
        engine.rollout(s0, np.repeat(d, 3, axis=1), t, DT, 20, hold=wl.HOLD)  # 3 steer channels
    r = engine.rollout(s0, d, t, DT, 0, hold=wl.HOLD)                       # zero steps: state passes through
    assert np.array_equal(r.state_end.cpu().numpy(), s0)
    # NaN propagates like the reference's numpy scalars (vx == 0), no error
    bad = s0.copy()
    bad[0, 0] = 0.0
    bad[1, 0] = 0.0
    bad[2, 0] = 0.0
    r = engine.rollout(bad, d, t, DT, 5, hold=wl.HOLD)
    out = r.state_end.cpu().numpy()
    assert np.isnan(out[0, 0]) and np.isfinite(out[:, 1:]).all()


def test_rollout_out_of_range_headings_take_checked_step(engine):
    """The fast-path kernel runs each RK4 step speculatively on branch-free heading trigonometry; headings beyond
    1e5 rad or stage increments |h*wz| > 2^-10 must be detected and repeated on the checked step."""
    B, N = 4096, 60
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    s0 = s0.copy()
    s0[7, 0:512:2] = 3.0e5 + np.arange(256)
    s0[2, 1:512:2] = 12.0
    s0[7, 1024:1100] = -2.5e5
    engine.set_params(_params())
    res = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=20)
    ref = c_oracle.rollout(s0, d, t, _c_params(), DT, N, hold=wl.HOLD, store_stride=20)
    e = rel_err(res.traj.cpu().numpy(), ref["traj"])
    mixed = np.zeros(B, dtype=bool)
    mixed[:512] = True
    mixed[1024:1100] = True
    assert e[:, :, ~mixed].max() < REL_TOL_F64
    # sin/cos of a 3e5 rad heading amplify a 1-ulp argument difference by ~3e5 ulp: conditioning of the
    # reference's own arithmetic, so the bound on those rows is looser by that factor
    assert e[:, :, mixed].max() < 1e-8, e[:, :, mixed].max()
    assert e[:, :8, mixed].max() < REL_TOL_F64          # body-frame states do not depend on the heading


def test_per_step_controls_pipeline_vs_c_oracle(engine):
    """hold = 1 on the front-steer fast path (the sampling-MPC shape): the kernel prepares the controls of step n + 1
    inside step n.  Same trajectories as the oracle, including a steer angle beyond the branch-free sincos range
    (library path), single-step and two-step launches (pipeline prologue / clamped look-ahead) and a resumed launch."""
    B, N = 640, 60
    rng = np.random.default_rng(3)
    s0, _, _ = wl.config2_rollouts(B=B, n_steps=N)
    d = rng.uniform(-0.1, 0.1, (N, 1, B))
    t = rng.uniform(-300.0, 300.0, (N, 1, B))
    d[7, 0, 5] = 2.0e5 + 0.3          # |delta| > 1e5: outside the fast range reduction
    d[N - 1, 0, 9] = -3.0e5 - 0.2     # ... in the last step, where the look-ahead index is clamped
    engine.set_params(_params())
    ref = c_oracle.rollout(s0, d, t, _c_params(), DT, N, hold=1, store_stride=1)
    got = engine.rollout(s0, d, t, DT, N, hold=1, store_stride=1)
    e = rel_err(got.traj.cpu().numpy(), ref["traj"])
    assert e.max() < REL_TOL_F64, e.max()
    # the FP32 twin takes the same pipelined path: inside its drift bound on the rollouts with ordinary steer angles
    g32 = engine.rollout(s0, d, t, DT, N, hold=1, dtype="f32")
    ok = np.ones(B, dtype=bool)
    ok[[5, 9]] = False
    e32 = rel_err(g32.state_end.cpu().numpy().astype(np.float64)[:10, ok], ref["state_end"][:10, ok])
    assert e32.max() < 2e-4, e32.max()
    for n1 in (1, 2, 31):             # split launches reproduce the single launch bit for bit
        a = engine.rollout(s0, d, t, DT, n1, hold=1, store_stride=1)
        b = engine.rollout(a.state_end, d, t, DT, N - n1, hold=1, store_stride=1, step0=n1)
        assert torch.equal(torch.cat([a.traj, b.traj]), got.traj)
        assert torch.equal(b.state_end, got.state_end)


def test_logging_rollout_with_stride_vs_c_oracle(engine):
    """want_aux with store_stride > 1 (the open-loop DataLog): the kernel takes the tabulated step between the stored
    steps and the closed-form logging step on them; trajectory, state_dot and the 18 outputs against the C oracle,
    front-steer and 4-channel layouts, and the assembled 45-column rows."""
    B, N, stride = 1000, 120, 10
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    engine.set_params(_params())
    ref = c_oracle.rollout(s0, d, t, _c_params(), DT, N, hold=wl.HOLD, store_stride=stride, want_aux=True)
    got = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=stride, want_aux=True)
    assert rel_err(got.traj.cpu().numpy(), ref["traj"]).max() < REL_TOL_F64
    assert rel_err(got.aux.cpu().numpy(), ref["aux"]).max() < REL_TOL_F64
    assert rel_err(got.state_end.cpu().numpy(), ref["state_end"]).max() < REL_TOL_F64
    z = np.zeros_like(d)
    d4, t4 = np.concatenate([d, d, z, z], axis=1), np.repeat(t, 4, axis=1)
    got4 = engine.rollout(s0, d4, t4, DT, N, hold=wl.HOLD, store_stride=stride, want_aux=True)
    assert rel_err(got4.aux.cpu().numpy(), ref["aux"]).max() < REL_TOL_F64


def test_state_broadcast_winner_record_and_per_call_friction(engine):
    """The resident sampling-MPC pieces: a [12,1] start state read with a zero batch stride gives the same bits as the
    materialised [12,B] copy; ``b200mp_mpc_winner_f64`` = lowest-index argmin + the winner's control sequence; the
    per-launch friction override equals the process-wide mode."""
    cfg = wl.config4_mpc(B=4096 + 37, n_steps=40)
    B, N = cfg["B"], cfg["n_steps"]
    engine.set_params(_params())
    d, t = engine.mpc_sample_controls(B, N, cfg["seed"])
    s0 = np.repeat(cfg["state0"][:, None], B, axis=1)
    kw = dict(hold=1, cost_ref=cfg["cost_ref"], w_u=cfg["w_u"], u_ref=cfg["u_ref"])
    full = engine.rollout(s0, d, t, DT, N, **kw)
    bc = engine.rollout(cfg["state0"][:, None], d, t, DT, N, state_broadcast=True, **kw)
    assert torch.equal(full.cost, bc.cost) and torch.equal(full.state_end, bc.state_end)
    # hold = 10 (the sliced / plain kernels) and a zero-step launch
    full10 = engine.rollout(s0, d, t, DT, N, hold=10)
    bc10 = engine.rollout(cfg["state0"][:, None], d, t, DT, N, hold=10, state_broadcast=True)
    assert torch.equal(full10.state_end, bc10.state_end)
    z = engine.rollout(cfg["state0"][:, None], d, t, DT, 0, hold=1, state_broadcast=True)
    assert np.array_equal(z.state_end.cpu().numpy(), s0)
    with pytest.raises(ValueError):
        engine.rollout(s0, d, t, DT, N, state_broadcast=True, **kw)
    cost = bc.cost.cpu().numpy()
    i = mpc_numpy.argmin_lowest(cost)
    rec = engine.mpc_winner(bc.cost, d, t, index_offset=1000).cpu().numpy()
    assert rec[0] == cost[i] and rec[1] == i + 1000
    assert np.array_equal(rec[2:2 + N], d[:, 0, i].cpu().numpy()) and np.array_equal(rec[2 + N:], t[:, 0, i].cpu().numpy())
    tie = bc.cost.clone()
    tie[7] = tie[3000] = -1.0                       # ties -> lowest index; NaN never wins
    tie[5] = float("nan")
    assert engine.mpc_winner(tie, d, t).cpu().numpy()[1] == 7
    none = engine.mpc_winner(torch.full((B,), float("inf"), dtype=torch.float64, device="cuda"), d, t).cpu().numpy()
    assert none[1] == -1 and np.isinf(none[0]) and not none[2:].any()
    # per-launch friction override == process-wide mode, both ways
    for mode in ("auto", "closed_form"):
        a = engine.rollout(s0, d, t, DT, N, hold=10, friction=mode)
        prev = engine.set_friction_mode(mode)
        b = engine.rollout(s0, d, t, DT, N, hold=10)
        engine.set_friction_mode(prev)
        assert torch.equal(a.state_end, b.state_end)
    with pytest.raises(ValueError):
        engine.rollout(s0, d, t, DT, N, friction="table")


def test_param_set_range_and_endstate_pipeline_segments(engine):
    """(i) per-rollout parameter-set indices outside the uploaded table are rejected on the host when they come
    as host data and clamped in the kernel when they come as a device tensor (never an out-of-bounds read);
    (ii) ``rollout_endstate_to_host`` with ``hold`` not dividing the chunk waits for the upload that covers the LAST
    control segment a launch reads (n_steps = 100, hold = 30 needs segment 3)."""
    B, N = 512, 100
    s0, _, _ = wl.config2_rollouts(B=B, n_steps=N)
    rng = np.random.default_rng(5)
    sets = np.stack([rng.uniform(8.0, 25.0, 3), rng.uniform(1.2, 1.9, 3), rng.uniform(0.3, 1.2, 3)], 1)
    p = VehicleParameters()
    for w in ("FL", "FR", "RL", "RR"):
        setattr(p, "B" + w, sets[:, 0]); setattr(p, "C" + w, sets[:, 1]); setattr(p, "D" + w, sets[:, 2])
    assert engine.set_params(p) == 3
    d = rng.uniform(-0.05, 0.05, (4, 1, B))
    t = rng.uniform(-200.0, 200.0, (4, 1, B))
    ps = rng.integers(0, 3, B).astype(np.int32)
    good = engine.rollout(s0, d, t, DT, N, hold=30, param_set=ps)
    bad = ps.copy()
    bad[7], bad[9] = 3, -1
    with pytest.raises(ValueError):
        engine.rollout(s0, d, t, DT, N, hold=30, param_set=bad)
    with pytest.raises(ValueError):
        engine.planar_model_batch(s0[:10], np.zeros((4, B)), None, np.zeros((4, B)), np.zeros(B), np.zeros(B), param_set=bad)
    # as a device tensor the indices are not read back: the kernel clamps them to [0, n_sets)
    clamped = ps.copy()
    clamped[7], clamped[9] = 2, 0
    got = engine.rollout(s0, d, t, DT, N, hold=30, param_set=torch.from_numpy(bad).cuda())
    want = engine.rollout(s0, d, t, DT, N, hold=30, param_set=clamped)
    assert torch.equal(got.state_end, want.state_end)
    keep = np.ones(B, dtype=bool)
    keep[[7, 9]] = False
    assert torch.equal(got.state_end[:, torch.from_numpy(keep).cuda()], good.state_end[:, torch.from_numpy(keep).cuda()])
    # (ii) end-state pipeline, hold = 30 (segments 0..3 for 100 steps), one chunk and several chunk sizes
    engine.set_params(_params())
    dev = engine.rollout(s0, d, t, DT, N, hold=30)
    hs, hd, ht = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (s0, d, t))
    end_host = torch.empty(12, B, dtype=torch.float64).pin_memory()
    for chunk in (100, 200, 60, 30):
        end_host.zero_()
        engine.rollout_endstate_to_host(hs, hd, ht, DT, N, 30, end_host, chunk_steps=chunk)
        assert torch.equal(end_host, dev.state_end.cpu()), chunk


def test_host_pipeline_slab_ring_and_pinned_buffers(engine):
    """``Engine.pinned_empty`` (NUMA-local when the host allows it, plain pinned memory otherwise) feeds the host-buffer
    pipeline; any number of device slabs gives the bytes of one device launch."""
    B, N = 4096, 120
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    engine.set_params(_params())
    dev = engine.rollout(s0, d, t, DT, N, hold=wl.HOLD, store_stride=1)
    hs, hd, ht = (torch.from_numpy(a).pin_memory() for a in (s0, d, t))
    out = engine.pinned_empty(N, 10, B)
    assert out.is_pinned() and out.shape == (N, 10, B) and hasattr(out, "numa_node")
    for n_slabs, chunk in ((2, 30), (3, 20), (5, 10), (4, 120)):
        out.zero_()
        end = engine.rollout_to_host(hs, hd, ht, DT, N, wl.HOLD, out, chunk_steps=chunk, n_slabs=n_slabs)
        assert torch.equal(out, dev.traj.cpu()) and torch.equal(end, dev.state_end), (n_slabs, chunk)
