"""Pinned host buffers for the host-buffer entry points (``Engine.rollout_to_host`` ...).

``pinned_empty`` returns a page-locked CPU tensor.  With ``numa_node >= 0`` its pages are first-touched while the calling
thread's memory policy is bound to that NUMA node (``set_mempolicy(MPOL_BIND)``) and then registered with
``cudaHostRegister`` -- i.e. the buffer lives on the node the GPU's PCIe root hangs off, so a device->host copy does not
cross the socket interconnect.  ``gpu_numa_node`` reads that node from sysfs.  Plumbing only: nothing here computes.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import torch

_NR_SET_MEMPOLICY = 238      # x86_64
_MPOL_DEFAULT, _MPOL_BIND = 0, 2


def gpu_numa_node(device: int) -> int:
    """NUMA node of the CUDA device's PCIe function (``-1``: unknown / single node)."""
    try:
        p = torch.cuda.get_device_properties(device)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return -1


def _set_mempolicy(mode: int, node: int = -1) -> bool:
    libc = ctypes.CDLL(None, use_errno=True)
    if mode == _MPOL_DEFAULT:
        return libc.syscall(_NR_SET_MEMPOLICY, _MPOL_DEFAULT, None, ctypes.c_ulong(0)) == 0
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(_NR_SET_MEMPOLICY, mode, ctypes.byref(mask), ctypes.c_ulong(64)) == 0


def pinned_empty(*shape, dtype=torch.float64, numa_node: int = -1) -> torch.Tensor:
    """Page-locked CPU tensor; ``numa_node >= 0`` places it on that node (falls back to ``pin_memory`` when the policy
    cannot be set, e.g. a container without the node in its ``mems``).  ``t.numa_node`` records where it was placed."""
    if numa_node is not None and numa_node >= 0 and os.name == "posix" and _set_mempolicy(_MPOL_BIND, numa_node):
        try:
            t = torch.empty(*shape, dtype=dtype)
            t.view(-1).zero_()                                    # first touch under the bound policy
        finally:
            _set_mempolicy(_MPOL_DEFAULT)
        rt = torch.cuda.cudart()
        err = rt.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        if int(err) == 0:
            ptr = t.data_ptr()
            weakref.finalize(t, lambda p=ptr: rt.cudaHostUnregister(p))
            t.numa_node = numa_node
            return t
    t = torch.empty(*shape, dtype=dtype).pin_memory()
    t.numa_node = -1
    return t
