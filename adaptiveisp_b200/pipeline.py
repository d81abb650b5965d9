"""Host <-> device staging around the ISP step.

The reference's training loop uploads the next image batch at the top of every iteration and
downloads the retouched batch at the bottom, both synchronously on the compute stream
(train.py:255 ``imgs.to(device)`` and :378-381 ``retouch.detach().cpu().numpy()``).  At B200 kernel
speeds those two PCIe transfers cost more than all ten filters forward + backward, so the caller-
facing loop overlaps them with compute on two copy streams:

    loop = HostStagedLoop(device)
    for (x,) in loop.stage(host_batches):        # batch k+1 uploads while batch k computes
        y = ...filters...(x)
        loop.fetch(y, pinned_host_out)           # batch k downloads while batch k+1 computes
    loop.drain()

Only CUDA streams/events and the caching allocator's ``record_stream`` are used -- no extra copies,
no threads.  Host tensors must be pinned for the copies to be asynchronous.
"""
from __future__ import annotations

from collections import deque
from typing import Iterable, Iterator, Sequence, Tuple

import torch


def _static_grads(modules, step_fn, example_inputs, dev, warmup, bucket=False):
    """Warm `step_fn` up on a side stream, then give every parameter that received a gradient a
    STATIC zero ``.grad`` tensor (parameters that never get one -- ``fc_mask.*`` -- keep ``None``, as in the
    reference).  Captured with these in place, every graph ACCUMULATES into the same tensors on
    replay, whichever slot it belongs to; callers clear them with ``zero_grad(set_to_none=False)``.
    ``bucket=True``: the static gradients are views of ONE flat buffer (``dist.GradBucket``), so that a
    data-parallel caller all-reduces them with a single NCCL call and no copies; returns the bucket."""
    for m in modules:
        m.zero_grad(set_to_none=True)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(max(1, warmup)):
            step_fn(*example_inputs)
    torch.cuda.current_stream(dev).wait_stream(side)
    if bucket:
        from .dist import GradBucket
        b = GradBucket([p for m in modules for p in m.parameters()], only_with_grad=True)
        b.zero()
        return b
    static = []
    for m in modules:
        for p in m.parameters():
            if p.grad is not None:
                p.grad = torch.zeros_like(p)
                static.append(p.grad)
    return static


class HostStagedLoop:
    def __init__(self, device, depth: int = 2):
        self.dev = torch.device(device)
        self.depth = max(1, int(depth))
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._pending = deque()
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def stage(self, host_batches: Iterable[Sequence[torch.Tensor]]) -> Iterator[Tuple[torch.Tensor, ...]]:
        """Yield device copies of each host batch (a sequence of pinned tensors), keeping
        ``depth - 1`` uploads in flight behind the batch being computed."""
        it = iter(host_batches)
        queue = deque()

        def issue() -> bool:
            try:
                host = next(it)
            except StopIteration:
                return False
            with torch.cuda.stream(self.s_in):
                dev = tuple(t.to(self.dev, non_blocking=True) for t in host)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            self.h2d_bytes += sum(t.numel() * t.element_size() for t in host)
            queue.append((dev, ev))
            return True

        for _ in range(self.depth):
            if not issue():
                break
        while queue:
            dev, ev = queue.popleft()
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_event(ev)
            for t in dev:
                t.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
            issue()
            yield dev

    def fetch(self, dev_tensor: torch.Tensor, host_tensor: torch.Tensor) -> None:
        """Asynchronous device -> pinned-host copy ordered after everything queued so far on the
        current stream; successive fetches are serialised on the copy-out stream."""
        cur = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.s_out.wait_event(ready)
        with torch.cuda.stream(self.s_out):
            host_tensor.copy_(dev_tensor, non_blocking=True)
            dev_tensor.record_stream(self.s_out)
            done = torch.cuda.Event()
            done.record(self.s_out)
        self.d2h_bytes += dev_tensor.numel() * dev_tensor.element_size()
        self._pending.append(done)
        while len(self._pending) > 8:
            self._pending.popleft().synchronize()

    def drain(self) -> None:
        """Block until every download has landed and make the compute stream wait for the copy
        streams (so that a following ``torch.cuda.synchronize`` / event covers them)."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_out)
        while self._pending:
            self._pending.popleft().synchronize()


class GraphedHostLoop:
    """CUDA-graph replay of a whole ISP step (forward AND backward, FC layers included) fed from
    pinned host memory, with double-buffered static input slots so that the upload of batch k+1 and
    the download of batch k-1's result overlap the replay of batch k.

    At B200 speed the eager class API is CPU-launch-bound (about 400 small launches per 10-filter
    step); one graph launch per step removes that, and is legal here because the per-step filter
    application has static shapes and no host-side decisions.

        loop = GraphedHostLoop(step_fn, example_inputs=(x_dev, feat_dev), modules=filters)
        loop.run(host_batches, host_out)        # host_batches: iterable of tuples of pinned tensors

    ``step_fn(*device_inputs)`` may call ``.backward()``; parameter ``.grad`` tensors become static and
    every replay (of any slot) ACCUMULATES into them, like an eager ``backward()`` without
    ``zero_grad`` (use ``zero_grad(set_to_none=False)`` between optimizer steps).  Returns nothing; the result of
    every step is copied asynchronously into ``host_out``: one pinned tensor reused for every step,
    or a sequence of pinned tensors, one per batch.
    """

    class _Slot:
        pass

    def __init__(self, step_fn, example_inputs, modules=(), slots: int = 2, warmup: int = 2):
        self.dev = example_inputs[0].device
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.slots = []
        self.static_grads = _static_grads(modules, step_fn, example_inputs, self.dev, warmup)
        for _ in range(max(1, slots)):
            s = GraphedHostLoop._Slot()
            s.inputs = tuple(t.clone() for t in example_inputs)
            s.graph = torch.cuda.CUDAGraph()
            for g in self.static_grads:      # same .grad state (allocated, zero) before EVERY capture
                g.zero_()
            with torch.cuda.graph(s.graph):
                s.out = step_fn(*s.inputs)
            s.ready = torch.cuda.Event()
            s.done = torch.cuda.Event()
            s.out_free = torch.cuda.Event()
            cur = torch.cuda.current_stream(self.dev)
            s.done.record(cur)
            s.out_free.record(cur)
            self.slots.append(s)
        torch.cuda.synchronize(self.dev)

    def run(self, host_batches, host_out=None) -> int:
        cur = torch.cuda.current_stream(self.dev)
        n = 0
        for k, host in enumerate(host_batches):
            s = self.slots[k % len(self.slots)]
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(s.done)            # the slot's previous replay has consumed its inputs
                for dst, src in zip(s.inputs, host):
                    dst.copy_(src, non_blocking=True)
                    self.h2d_bytes += src.numel() * src.element_size()
                s.ready.record(self.s_in)
            cur.wait_event(s.ready)
            cur.wait_event(s.out_free)                  # its previous result has left for the host
            s.graph.replay()
            s.done.record(cur)
            if host_out is not None:
                dst = host_out[k] if isinstance(host_out, (list, tuple)) else host_out
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(s.done)
                    dst.copy_(s.out, non_blocking=True)
                    s.out_free.record(self.s_out)
                self.d2h_bytes += s.out.numel() * s.out.element_size()
            n += 1
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_out)
        return n


class GraphedStep:
    """One CUDA graph around a device-resident step (e.g. ``Agent.forward`` + ``backward``).

    On a B200 the eager Agent step is CPU-launch-bound (about 7 ms of launches for 1.5-3.5 ms of GPU
    work, ``scripts/agent_step_timing.py``); replaying it as a graph removes that.  Shapes and the
    set of modules must stay fixed; ``step_fn`` may call ``.backward()`` (parameter ``.grad`` tensors
    become static and each replay accumulates into them: clear them with ``zero_grad(set_to_none=False)``).

        g = GraphedStep(step_fn, example_inputs=(x, z, states), modules=[agent])
        out = g(x_new, z_new, states_new)       # copies into the static inputs, replays, returns static outputs
    """

    def __init__(self, step_fn, example_inputs, modules=(), warmup: int = 2, grad_bucket: bool = False):
        dev = example_inputs[0].device
        self.bucket = None
        if grad_bucket:
            self.bucket = _static_grads(modules, step_fn, example_inputs, dev, warmup, bucket=True)
            self.static_grads = [self.bucket.flat]
        else:
            self.static_grads = _static_grads(modules, step_fn, example_inputs, dev, warmup)
        self.inputs = tuple(t.clone() for t in example_inputs)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = step_fn(*self.inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
