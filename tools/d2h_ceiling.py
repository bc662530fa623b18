#!/usr/bin/env python
"""What the box allows for the trajectory read-back: device->host copy ceiling per GPU and for all GPUs at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_ceiling.py

Every rank copies a 262 MB device slab (one 50-step chunk of config 2) into pinned host memory with
cudaMemcpyAsync, 40 times back to back, (a) one rank at a time, (b) all ranks at once, for pinned buffers allocated
(1) as torch does (cudaHostAlloc from wherever the thread runs) and (2) after binding the allocating thread's memory
policy to the NUMA node the GPU hangs off (set_mempolicy(MPOL_BIND), first touch by this thread).  Also records the
topology facts that explain the numbers: NUMA nodes visible, the GPU's node, CPU affinity.  Rank 0 prints one JSON object.
"""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist


def numa_facts(dev_index):
    out = {"nodes": [], "gpu_numa_node": None, "affinity_cpus": len(os.sched_getaffinity(0))}
    base = "/sys/devices/system/node"
    try:
        for n in sorted(os.listdir(base)):
            if n.startswith("node") and n[4:].isdigit():
                out["nodes"].append({"node": int(n[4:]), "cpulist": open(f"{base}/{n}/cpulist").read().strip()})
    except Exception as exc:  # noqa: BLE001
        out["nodes_error"] = str(exc)
    try:
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devid = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        out["gpu_pci"] = f"{dom:04x}:{bus:02x}:{devid:02x}.0"
        out["gpu_numa_node"] = int(open(path).read().strip())
    except Exception as exc:  # noqa: BLE001
        out["gpu_numa_error"] = str(exc)
    try:
        out["mems_allowed"] = [ln.split(":")[1].strip() for ln in open("/proc/self/status") if ln.startswith("Mems_allowed_list")][0]
    except Exception:
        pass
    return out


def bind_memory(node):
    """set_mempolicy(MPOL_BIND, {node}) for the calling thread; returns True on success."""
    if node is None or node < 0:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, 2, ctypes.byref(mask), ctypes.c_ulong(64))     # __NR_set_mempolicy, MPOL_BIND
    return rc == 0


def unbind_memory():
    libc = ctypes.CDLL(None, use_errno=True)
    libc.syscall(238, 0, None, ctypes.c_ulong(0))                         # MPOL_DEFAULT


def copy_rate(dst, src, reps, stream):
    with torch.cuda.stream(stream):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        stream.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        stream.synchronize()
    return src.numel() * src.element_size() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    facts = numa_facts(local)
    n = 50 * 10 * 65536                               # doubles in one 50-step chunk of config 2 (262 MB)
    src = torch.empty(n, dtype=torch.float64, device=dev).normal_()
    stream = torch.cuda.Stream(dev)
    res = {}
    bufs = {"default": torch.empty(n, dtype=torch.float64).pin_memory()}
    bound = bind_memory(facts.get("gpu_numa_node"))
    facts["mempolicy_bind_ok"] = bound
    if bound:
        try:
            b = torch.empty(n, dtype=torch.float64)
            b.zero_()                                 # first touch under the bound policy
            bufs["numa_local"] = b.pin_memory() if False else b
            torch.cuda.cudart().cudaHostRegister(b.data_ptr(), b.numel() * 8, 0)
        except Exception as exc:  # noqa: BLE001
            facts["numa_local_error"] = str(exc)
            bufs.pop("numa_local", None)
        unbind_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for name, dst in bufs.items():
        alone = None
        for r in range(world):                        # one rank at a time
            barrier()
            if r == rank:
                alone = copy_rate(dst, src, 20, stream)
        barrier()
        together = copy_rate(dst, src, 40, stream)    # everybody at once
        barrier()
        h2d = copy_rate(src, dst, 20, stream)         # the other direction, all at once
        barrier()
        res[name] = {"d2h_alone_gbs": alone, "d2h_all_gbs": together, "h2d_all_gbs": h2d}
    allres = [None] * world
    if world > 1:
        dist.all_gather_object(allres, {"rank": rank, "facts": facts, "rates": res})
    else:
        allres = [{"rank": rank, "facts": facts, "rates": res}]
    if rank == 0:
        summary = {"world": world, "chunk_bytes": n * 8}
        for name in bufs:
            rows = [a["rates"].get(name) for a in allres if a["rates"].get(name)]
            if rows:
                summary[name] = {"d2h_alone_gbs_per_gpu": [round(r["d2h_alone_gbs"], 1) for r in rows],
                                 "d2h_all_gbs_per_gpu": [round(r["d2h_all_gbs"], 1) for r in rows],
                                 "d2h_all_gbs_aggregate": round(sum(r["d2h_all_gbs"] for r in rows), 1),
                                 "h2d_all_gbs_aggregate": round(sum(r["h2d_all_gbs"] for r in rows), 1)}
        summary["per_rank_facts"] = [a["facts"] for a in allres]
        print(json.dumps(summary))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
