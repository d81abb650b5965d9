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


class GradBucket:
    """ONE persistent flat fp32 buffer whose views ARE the parameters' ``.grad`` tensors.

    The reference's training step ends with ``loss.backward()`` -> (data-parallel) gradient all-reduce ->
    ``clip_grad_norm`` -> Adam (train.py:341-349).  With the gradients living in one buffer the
    all-reduce is a single NCCL call on memory that backward wrote in place -- no ``cat`` before and no
    copy-back after -- the clip is one norm over the same buffer, and a CUDA graph that captured the
    backward keeps accumulating into it (``pipeline.GraphedStep`` uses static ``.grad`` tensors anyway).

    ``params``: the parameters that receive a gradient.  Parameters that never do (``fc_mask.*``) must be
    left out -- pass ``only_with_grad=True`` after one warm-up backward to select exactly those whose
    ``.grad`` is not ``None`` (their current values are carried over).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], only_with_grad: bool = False):
        self.params = [p for p in params if p.requires_grad and (p.grad is not None or not only_with_grad)]
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("GradBucket parameters must share device and dtype")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """True while every ``.grad`` still aliases the bucket (``zero_grad(set_to_none=True)`` detaches)."""
        base = self.flat.data_ptr()
        end = base + self.nbytes
        return all(p.grad is not None and base <= p.grad.data_ptr() < end for p in self.params)

    @torch.no_grad()
    def allreduce(self, average: bool = True, group=None, async_op: bool = False):
        """Sum (or average) the bucket across ranks in place with one all-reduce; returns the work handle
        when ``async_op`` (the division then is the caller's job after ``wait()``)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        if average:
            self.flat.div_(dist.get_world_size(group))
        return None

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """``torch.nn.utils.clip_grad_norm_`` (train.py:345-346) over the flat buffer: one norm, one scale,
        no host sync."""
        total = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total


@torch.no_grad()
def allreduce_grads(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """Sum (or average) the existing ``.grad`` tensors across ranks with one flat all-reduce.

    Parameters whose grad is ``None`` on this rank (e.g. ``fc_mask.*``, which never receive a
    gradient) are skipped -- they are ``None`` on every rank.  Unselected filters carry exact-zero
    grads (not ``None``), so the flat layout is identical on all ranks.  Returns the bucket size in
    bytes (0 when not distributed).

    The gradients are moved into a persistent :class:`GradBucket` on the first call (cached on the
    first parameter): from then on backward writes straight into the flat buffer and every later call
    is one NCCL launch with no gather / scatter copies.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    if dist.get_world_size(group) == 1:
        return 0
    plist = [p for p in params if p.grad is not None]
    if not plist:
        return 0
    bucket = getattr(plist[0], "_aisp_grad_bucket", None)
    if bucket is None or len(bucket.params) != len(plist) or any(a is not b for a, b in zip(bucket.params, plist)) \
            or not bucket.attached():
        bucket = GradBucket(plist, only_with_grad=True)
        plist[0]._aisp_grad_bucket = bucket
    bucket.allreduce(average=average, group=group)
    return bucket.nbytes


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
