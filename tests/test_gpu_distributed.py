"""Multi-GPU path on the GPU box.

With one GPU the ranks are emulated in ONE process (every "rank" runs its shard on cuda:0 in turn and the
pairs are combined by the same host function the collective feeds), which checks the property the real
multi-GPU run relies on: results do not depend on how many ranks the batch is sharded over.  With >= 2
GPUs the real NCCL path is launched (tools/dist_check.py under torch.distributed.run).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import c_oracle
from python_motionplanning_b200 import VehicleParameters, distributed as D, workloads as wl

pytestmark = pytest.mark.gpu
OFF, RAD = list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII)


def _mpc_shard(engine, cfg, lo, hi, hold=1):
    n_seg = -(-cfg["n_steps"] // hold)
    d, t = engine.mpc_sample_controls(hi - lo, n_seg, cfg["seed"], rollout0=lo, delta_mean=cfg["delta_mean"],
                                      delta_sigma=cfg["delta_sigma"], delta_clip=cfg["delta_clip"],
                                      torque_mean=cfg["torque_mean"], torque_sigma=cfg["torque_sigma"])
    s0 = engine.dev(cfg["state0"]).reshape(12, 1).expand(12, hi - lo).contiguous()
    res = engine.rollout(s0, d, t, wl.DT, cfg["n_steps"], hold=hold, cost_ref=cfg["cost_ref"], w_u=cfg["w_u"], u_ref=cfg["u_ref"])
    mn, ix = engine.argmin(res.cost, index_offset=lo)
    return float(mn.item()), int(ix.item()), res.cost, d, t


def test_mpc_result_independent_of_rank_count(engine):
    cfg = wl.config4_mpc(B=65536, n_steps=100)
    p = VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    engine.set_params(p)
    mn1, ix1, cost1, d1, _ = _mpc_shard(engine, cfg, 0, cfg["B"])
    for ws in (2, 3, 8):
        pairs, costs = [], []
        for r in range(ws):
            lo, hi = D.shard_range(cfg["B"], r, ws)
            mn, ix, cost, _, _ = _mpc_shard(engine, cfg, lo, hi)
            pairs.append((mn, ix))
            costs.append(cost)
        cost, idx, owner = D.pick_winner(pairs)
        assert (cost, idx) == (mn1, ix1)
        assert D.shard_range(cfg["B"], owner, ws)[0] <= idx < D.shard_range(cfg["B"], owner, ws)[1]
        assert torch.equal(torch.cat(costs), cost1)          # bit-identical per-rollout costs
    # the single-process mpc_plan entry point agrees
    plan = D.mpc_plan(engine, cfg)
    assert plan["index"] == ix1 and plan["cost"] == mn1 and torch.equal(plan["delta"], d1[:, 0, ix1].cpu())


def test_collision_sharded_equals_unsharded(engine):
    w = wl.config3_lattice(P=1000, M=4000)
    full = engine.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD).cpu().numpy()
    for ws in (2, 7):
        parts = []
        for r in range(ws):
            lo, hi = D.shard_range(1000, r, ws)
            parts.append(engine.collision_check_batch(w["px"][lo:hi], w["py"][lo:hi], w["pyaw"][lo:hi], w["obstacles"], OFF, RAD).cpu().numpy())
        assert np.array_equal(np.concatenate(parts), full)
    free, best = D.collision_select_sharded(engine, w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD, w["goal"][:2], w["weight"])
    ref, _, _ = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], OFF, RAD)
    assert np.array_equal(free.cpu().numpy().astype(bool), ref)


def test_sharded_lattice_and_tracking_single_process(engine):
    """World size 1 paths of plan_lattice_sharded / track_sharded equal the plain engine calls; emulated shards of the
    fleet (whole waypoint sets per rank) reproduce the unsharded end states bit for bit."""
    p = VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    engine.set_params(p)
    rng = np.random.default_rng(9)
    goals = np.stack([rng.uniform(22, 38, 200), rng.uniform(-6, 6, 200), rng.uniform(-0.3, 0.3, 200)])
    w = wl.config3_lattice(P=8, M=2000)
    obs = w["obstacles"] * 0.6 - 10.0
    b1, out1 = engine.plan_lattice(goals, (12.0, -7.0, 0.4), obs, OFF, RAD, (45.0, 10.0), w["weight"])
    b2, free2, end2 = D.plan_lattice_sharded(engine, goals, (12.0, -7.0, 0.4), obs, OFF, RAD, (45.0, 10.0), w["weight"])
    assert b1 == b2 and torch.equal(free2, out1["free"]) and torch.equal(end2, out1["end_xy"])
    st0, wps = wl.tracking_fleet(V=1000, n_sets=5)
    one = engine.track_closed_loop(st0, wps, wl.DT, 40, vehicles_per_set=200)
    res, (lo, hi), full = D.track_sharded(engine, st0, wps, wl.DT, 40, gather=True)
    assert (lo, hi) == (0, 1000) and torch.equal(full, one.state_end)
    for ws in (2, 3):
        parts = []
        for r in range(ws):
            s_lo, s_hi = D.shard_range(5, r, ws)
            parts.append(engine.track_closed_loop(st0[:, s_lo * 200:s_hi * 200], wps[s_lo:s_hi], wl.DT, 40, vehicles_per_set=200).state_end)
        assert torch.equal(torch.cat(parts, dim=1), one.state_end)


def test_nccl_two_ranks_when_available():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CHECK_OK" in out.stdout
