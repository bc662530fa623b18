#!/usr/bin/env python
"""Run under torchrun on >= 2 GPUs: the NCCL path of distributed.mpc_plan / collision_select_sharded /
plan_lattice_sharded / track_sharded,
checked against the unsharded single-GPU result on every rank.  Prints DIST_CHECK_OK from rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import distributed as D, workloads as wl  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = mp.Engine(local)
    p = mp.VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    eng.set_params(p)
    cfg = wl.config4_mpc(B=1 << 18, n_steps=100)
    plan = D.mpc_plan(eng, cfg)
    # unsharded reference on this rank's own GPU
    d, t = eng.mpc_sample_controls(cfg["B"], 100, cfg["seed"])
    s0 = eng.dev(cfg["state0"]).reshape(12, 1).expand(12, cfg["B"]).contiguous()
    res = eng.rollout(s0, d, t, wl.DT, 100, hold=1, cost_ref=cfg["cost_ref"], w_u=cfg["w_u"], u_ref=cfg["u_ref"])
    mn, ix = eng.argmin(res.cost)
    assert plan["index"] == int(ix.item()) and plan["cost"] == float(mn.item()), (plan["index"], int(ix.item()))
    assert torch.equal(plan["delta"], d[:, 0, plan["index"]].cpu()) and torch.equal(plan["torque"], t[:, 0, plan["index"]].cpu())
    lo, hi = plan["shard"]
    assert torch.equal(plan["local_cost"], res.cost[lo:hi])
    w = wl.config3_lattice(P=4096, M=10000)
    free, best = D.collision_select_sharded(eng, w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"],
                                            w["goal"][:2], w["weight"])
    full = eng.collision_check_batch(w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"])
    assert torch.equal(free, full)
    best1 = eng.select_best_path_index_batch(w["px"][:, -1].copy(), w["py"][:, -1].copy(), full, w["goal"][:2], w["weight"])
    assert best == best1
    # sharded lattice planning (optimise -> sample -> check -> select) equals the single-GPU pipeline
    rng = np.random.default_rng(9)
    Pg = 1001
    gt = rng.uniform(-0.3, 0.3, Pg)
    goals = np.stack([rng.uniform(22, 38, Pg), rng.uniform(-6, 6, Pg), gt])
    ego = (12.0, -7.0, 0.4)
    obs = w["obstacles"] * 0.6 - 10.0
    b_sh, free_sh, end_sh = D.plan_lattice_sharded(eng, goals, ego, obs, w["offsets"], w["radii"], (45.0, 10.0), w["weight"])
    b_1, out1 = eng.plan_lattice(goals, ego, obs, w["offsets"], w["radii"], (45.0, 10.0), w["weight"])
    assert b_sh == b_1 and torch.equal(free_sh, out1["free"]) and torch.equal(end_sh, out1["end_xy"])
    # closed-loop fleet sharded by waypoint set equals the single-GPU run, bit for bit
    st0, wps = wl.tracking_fleet(V=4096, n_sets=5)
    res, (vlo, vhi), full_end = D.track_sharded(eng, st0, wps, wl.DT, 60, gather=True)
    one = eng.track_closed_loop(st0, wps, wl.DT, 60, vehicles_per_set=-(-4096 // 5))
    assert torch.equal(full_end, one.state_end)
    assert res is None or torch.equal(res.state_end, one.state_end[:, vlo:vhi])
    dist.barrier()
    if rank == 0:
        print(f"DIST_CHECK_OK world={dist.get_world_size()} mpc index={plan['index']} cost={plan['cost']:.6e} best path={best}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
