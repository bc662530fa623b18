"""world_size-2 ``gloo`` tests (CPU) of the multi-GPU glue: shard -> gather (cost, index) pairs ->
replicated lowest-index argmin -> broadcast of the winner; gather of per-path flags.

The exchange code is device-agnostic; on the GPU box the same functions run over NCCL.  The per-shard
values combined here come from the oracle (this is a test), never from a product CPU path.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from python_motionplanning_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, cost, flags, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n = len(cost)
        lo, hi = D.shard_range(n, rank, ws)
        local = torch.from_numpy(cost[lo:hi])
        # local lowest-index argmin with NaN = +inf (what b200mp_argmin_f64 returns on the GPU)
        c = torch.where(torch.isnan(local), torch.full_like(local, float("inf")), local)
        if torch.isfinite(c).any():
            i = int(torch.argmin(c))
            mn, ix = c[i].reshape(1), torch.tensor([lo + i])
        else:
            mn, ix = torch.tensor([float("inf")], dtype=torch.float64), torch.tensor([-1])
        best_cost, best_idx, owner = D.global_argmin(mn, ix)
        # broadcast of the winner's "control sequence" from its owner
        seq = torch.full((5,), float(rank + 1) * 100 + best_idx) if owner == rank else torch.zeros(5)
        D.broadcast_from(seq, owner)
        # the MPC exchange proper: one all-gather of per-rank winner records (cost, index, control sequence)
        n_seg = 5
        gcol = torch.arange(lo, hi, dtype=torch.float64)
        dl = torch.arange(n_seg, dtype=torch.float64).reshape(n_seg, 1) * 1000 + gcol            # value = 1000 seg + global index
        gw = D.gather_winner(mn.to(torch.float64), ix, dl, -dl, lo)
        assert gw[:3] == (best_cost, best_idx, owner)
        if best_idx >= 0:
            want = torch.arange(n_seg, dtype=torch.float64) * 1000 + best_idx
            assert torch.equal(gw[3], want) and torch.equal(gw[4], -want)
        fl = D.gather_flags(torch.from_numpy(flags[slice(*D.shard_range(len(flags), rank, ws))].copy()), len(flags))
        q.put((rank, best_cost, best_idx, owner, seq.tolist(), fl.numpy().tolist()))
    finally:
        dist.destroy_process_group()


def _run(cost, flags, ws=2):
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, cost, flags, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.parametrize("case", ["random", "tie_across_ranks", "nan_and_inf", "all_inf"])
def test_sharded_argmin_and_flag_gather(case):
    rng = np.random.default_rng(1)
    n = 1001
    cost = rng.uniform(1, 2, n)
    if case == "tie_across_ranks":
        cost[[100, 700]] = 0.5            # same minimum on both shards -> lowest global index wins
    elif case == "nan_and_inf":
        cost[3] = np.nan
        cost[600] = -np.inf if False else 0.25
        cost[10] = np.inf
    elif case == "all_inf":
        cost[:] = np.inf
    flags = (rng.uniform(size=777) < 0.4).astype(np.uint8)
    out = _run(cost, flags)
    c = np.where(np.isnan(cost), np.inf, cost)
    want_idx = int(np.argmin(c)) if np.isfinite(c).any() else -1
    want_owner = -1 if want_idx < 0 else (0 if want_idx < D.shard_range(n, 0, 2)[1] else 1)
    for rank, best_cost, best_idx, owner, seq, fl in out:
        assert best_idx == want_idx and owner == want_owner
        if want_idx >= 0:
            assert best_cost == c[want_idx]
            assert seq == [float(want_owner + 1) * 100 + want_idx] * 5      # every rank holds the owner's sequence
        assert fl == flags.tolist()


def test_single_process_paths():
    mn, ix = torch.tensor([0.75], dtype=torch.float64), torch.tensor([12])
    assert D.global_argmin(mn, ix) == (0.75, 12, 0)
    t = torch.arange(4.0)
    assert D.broadcast_from(t, 0) is t
    f = torch.tensor([1, 0, 1], dtype=torch.uint8)
    assert D.gather_flags(f, 3) is f
    assert D.world() == (0, 1)


def _gs_worker(rank, ws, port, full, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n = full.shape[-1]
        lo, hi = D.shard_range(n, rank, ws)
        got = D.gather_shards(torch.from_numpy(full[..., lo:hi].copy()), n)
        q.put((rank, got.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,n", [(2, 7), (2, 1), (3, 10)])
def test_gather_shards_float_rows(ws, n):
    """Path end points [2, P] split by columns across ranks come back in order on every rank (uneven shards too)."""
    rng = np.random.default_rng(4)
    full = rng.normal(size=(2, n))
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gs_worker, args=(r, ws, port, full, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, got in out:
        assert np.array_equal(got, full)
