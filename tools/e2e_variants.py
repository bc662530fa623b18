#!/usr/bin/env python
"""End-to-end (host buffers) config-2 rollouts on N GPUs under different read-back pipelines -- which knob moves it?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 tools/e2e_variants.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

B, N, HOLD = 65536, 500, 10


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = mp.Engine(local)
    p = mp.VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0
    eng.set_params(p)
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N, seed=wl.SEED + rank)
    hs, hd, ht = (torch.from_numpy(a).pin_memory() for a in (s0, d, t))
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(name, traj_host, **kw):
        for _ in range(2):
            eng.rollout_to_host(hs, hd, ht, wl.DT, N, HOLD, traj_host, **kw)
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            eng.rollout_to_host(hs, hd, ht, wl.DT, N, HOLD, traj_host, **kw)
        barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        out[name] = {"steps_per_s": world * B * N * reps / float(el.item()), "d2h_gbs_aggregate": world * N * 10 * B * 8 * reps / float(el.item()) / 1e9}

    th = torch.empty(N, 10, B, dtype=torch.float64).pin_memory()
    run("torch_pinned_2slabs_chunk50", th, chunk_steps=50, n_slabs=2)
    run("torch_pinned_4slabs_chunk50", th, chunk_steps=50, n_slabs=4)
    run("torch_pinned_4slabs_chunk20", th, chunk_steps=20, n_slabs=4)
    run("torch_pinned_2slabs_chunk100", th, chunk_steps=100, n_slabs=2)
    del th
    tn = eng.pinned_empty(N, 10, B)
    out["numa_node_of_buffer"] = int(getattr(tn, "numa_node", -1))
    run("numa_local_2slabs_chunk50", tn, chunk_steps=50, n_slabs=2)
    run("numa_local_4slabs_chunk50", tn, chunk_steps=50, n_slabs=4)
    if rank == 0:
        print(json.dumps({"world": world, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
