# SPDX-License-Identifier: Apache-2.0
"""NUMA placement of the host side of a GPU handle.

The host-buffer step is PCIe-bound: pinned buffers that live on the socket the GPU is NOT
attached to cross the inter-socket link on every step. ``bind_to_gpu_node`` pins the calling thread (and
therefore the pinned allocations it makes afterwards) to the CPUs of the GPU's NUMA node.
"""
import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device: int = 0) -> Optional[int]:
    """NUMA node of CUDA device ``device`` (sysfs), None when unknown or the machine has one node."""
    path = None
    try:
        import torch

        p = torch.cuda.get_device_properties(device)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
    except Exception:
        try:
            import pynvml

            pynvml.nvmlInit()
            bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
            path = f"/sys/bus/pci/devices/{bus_id[-12:].lower()}/numa_node"
        except Exception:
            return None
    try:
        with open(path) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device: int = 0) -> Optional[List[int]]:
    """Restrict the calling thread to the CPUs of the GPU's NUMA node; returns the previous affinity (pass it
    to ``os.sched_setaffinity(0, ...)`` to undo) or None when nothing was changed."""
    node = gpu_numa_node(device)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        previous = sorted(os.sched_getaffinity(0))
        allowed = sorted(set(cpus) & set(previous))
        if not allowed or allowed == previous:
            return None
        os.sched_setaffinity(0, allowed)
        return previous
    except Exception:
        return None
