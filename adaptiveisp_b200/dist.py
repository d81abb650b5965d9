"""Data-parallel plumbing for callers of the ISP path (one process per GPU, torch.distributed).

The ISP kernels themselves need no collective: every filter is per-image, so the batch is sharded
by sample and per-sample parameter gradients stay local.  What a training caller
needs is ONE all-reduce per optimizer step over the actor + critic gradients (7.18 M + 1.22 M
parameters = 33.6 MB fp32 in the reference's configuration), which is launch-latency- not
bandwidth-bound on NVLink 5 / NVSwitch; it is therefore issued as a single flat bucket.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range [lo, hi) of rank `rank`: sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _grads(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    return [p.grad for p in params if p.grad is not None]


@torch.no_grad()
def allreduce_grads(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """Sum (or average) the existing ``.grad`` tensors across ranks with one flat all-reduce.

    Parameters whose grad is ``None`` on this rank (e.g. ``fc_mask.*``, which never receive a
    gradient) are skipped -- they are ``None`` on every rank.  Unselected filters carry exact-zero
    grads (not ``None``), so the flat layout is identical on all ranks.  Returns the bucket size in
    bytes (0 when not distributed).
    """
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    grads = _grads(params)
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


# ---------------------------------------------------------------------------------------------
# NUMA placement of a rank's pinned host buffers.
# With one process per GPU every rank streams ~400 MB per step through pinned host memory
# (train.py:255, :378-381).  Pinned pages land on the NUMA node of the allocating thread; if that
# is not the node the GPU's PCIe root hangs off, every copy crosses the socket interconnect and the
# ranks of a node contend for it.  Pinning the process to the GPU's node before the buffers are
# allocated keeps each rank's traffic on its own memory controllers.
# ---------------------------------------------------------------------------------------------
def gpu_numa_node(device_index: int) -> int:
    """NUMA node of the GPU's PCIe function from sysfs, or -1 when unknown (single-node hosts, VMs)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return -1


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs of the GPU's NUMA node (so that pinned buffers allocated
    afterwards are node-local).  Returns ``(node, previous_affinity)``; ``node == -1`` means nothing
    was changed.  Undo with ``os.sched_setaffinity(0, previous_affinity)``."""
    import os
    prev = os.sched_getaffinity(0)
    node = gpu_numa_node(device_index)
    if node < 0:
        return -1, prev
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set(_parse_cpulist(f.read())) & prev
        if not cpus:
            return -1, prev
        os.sched_setaffinity(0, cpus)
        return node, prev
    except Exception:
        return -1, prev
