"""GPU-resident replay pool: the image pool of the reference's ``ReplayMemory`` kept in HBM.

The reference stores its pool of partially retouched images as NumPy arrays on the host
(replay_memory.py:100-221): every training iteration downloads the retouched batch
(train.py:378-381 ``retouch.detach().cpu().numpy()``) and uploads the next one -- from float64
(replay_memory.py:9-15, train.py:255).  At B200 kernel speeds those two PCIe transfers cost more
than the ten ISP filters forward and backward.  Here the pixel data and the agent states stay on the
device: a pool of 128 x 3 x 512 x 512 fp32 images is 403 MB of the 180 GB of HBM; only brand-new
images cross PCIe once, when they enter the pool.  Labels / paths / shapes (what the detector loss
needs) stay host-side Python objects addressed by slot.

Pool semantics follow the reference: batches are drawn uniformly at random without replacement from
the unfinished records (``get_next_fake_batch``: finished ones -- ``state[STATE_STOPPED_DIM] == 1`` --
are dropped when met); returned records re-enter unless their trajectory is over-long and loses the
``over_length_keep_prob`` coin flip (``replace_memory``); the pool is topped up with fresh images.

    pool = DeviceReplayPool(cfg, (3, 512, 512), device, fetch_fresh)   # fetch_fresh(n) -> (imgs[n,3,H,W], metas[n])
    batch = pool.get_batch(B)               # batch.images [B,3,H,W] (device), batch.states, batch.slots, batch.meta
    ...agent step...
    pool.put_back(batch.slots, retouched, new_states)
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

STATE_STOPPED_DIM, STATE_STEP_DIM = 1, 2   # util.py:15-18


@dataclass
class PoolBatch:
    images: torch.Tensor      # [B,3,H,W] fp32, on the pool's device (a gathered copy: safe to overwrite)
    states: torch.Tensor      # [B,S]
    slots: List[int]          # pool slots the samples came from
    meta: List[Any]           # per-sample host-side record (label, path, shape, ...)


class DeviceReplayPool:
    def __init__(self, cfg, image_shape: Tuple[int, int, int], device,
                 fetch_fresh: Callable[[int], Tuple[torch.Tensor, Sequence[Any]]],
                 capacity: Optional[int] = None, fetch_batch: int = 16, rng: Optional[random.Random] = None):
        self.cfg = cfg
        self.device = torch.device(device)
        self.capacity = int(capacity if capacity is not None else cfg.replay_memory_size)
        self.fetch_fresh = fetch_fresh
        self.fetch_batch = int(fetch_batch)
        self.rng = rng or random.Random()
        c, h, w = image_shape
        self.images = torch.empty((self.capacity, c, h, w), dtype=torch.float32, device=self.device)
        self.states = torch.zeros((self.capacity, cfg.num_state_dim), dtype=torch.float32, device=self.device)
        # host mirrors of the two state fields the pool logic branches on (a few bytes per record)
        self._stopped = [0.0] * self.capacity
        self._steps = [0.0] * self.capacity
        self.meta: List[Any] = [None] * self.capacity
        self._live: List[int] = []                       # slots holding a record
        self._free: List[int] = list(range(self.capacity))
        self.h2d_bytes = 0
        # fresh frames are uploaded on a side stream into a staging buffer: the PCIe copy overlaps whatever
        # the compute stream is still running (the step that was just enqueued) and only the device-side
        # scatter into the pool waits for it
        self._h2d = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.fill()

    # ------------------------------------------------------------------------------------------
    def _index(self, values: Sequence[int]) -> torch.Tensor:
        """Slot / row indices as a device tensor WITHOUT a blocking copy: torch.tensor(list, device=cuda)
        goes through pageable memory and stalls the host until the stream drains, which would serialise
        every training iteration; a pinned staging copy keeps the enqueue asynchronous."""
        t = torch.tensor(list(values), dtype=torch.long)
        if self.device.type != "cuda":
            return t
        return t.pin_memory().to(self.device, non_blocking=True)

    def __len__(self) -> int:
        return len(self._live)

    def fill(self) -> None:
        """replay_memory.py:126-140: top the pool up with brand-new images (initial states = 0)."""
        while self._free:
            n = min(self.fetch_batch, len(self._free))
            imgs, metas = self.fetch_fresh(n)
            if imgs.shape[0] != n or len(metas) != n:
                raise ValueError("fetch_fresh(n) must return n images and n meta records")
            slots = [self._free.pop() for _ in range(n)]
            idx = self._index(slots)
            if self._h2d is not None:
                with torch.cuda.stream(self._h2d):
                    src = imgs.to(self.device, dtype=torch.float32, non_blocking=True)
                    landed = torch.cuda.Event()
                    landed.record(self._h2d)
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(landed)
                src.record_stream(cur)
            else:
                src = imgs.to(self.device, dtype=torch.float32)
            self.h2d_bytes += imgs.numel() * imgs.element_size()
            self.images.index_copy_(0, idx, src)
            self.states.index_fill_(0, idx, 0.0)
            for s, m in zip(slots, metas):
                self.meta[s] = m
                self._stopped[s] = 0.0
                self._steps[s] = 0.0
            self._live.extend(slots)

    def get_batch(self, batch_size: int) -> PoolBatch:
        """replay_memory.py:208-221: a random batch of unfinished records, removed from the pool."""
        if batch_size > self.capacity:
            raise ValueError("batch larger than the pool")
        chosen: List[int] = []
        while len(chosen) < batch_size:
            if not self._live:
                # every record that is not in `chosen` is checked out by a caller who has not called
                # put_back / discard yet: nothing to draw from and nothing to refill
                if not self._free:
                    for s in chosen:
                        self._live.append(s)
                    raise RuntimeError(
                        f"replay pool exhausted: {self.capacity - len(chosen)} of {self.capacity} records are "
                        f"checked out (get_batch without put_back/discard); cannot draw {batch_size}")
                self.fill()
            self.rng.shuffle(self._live)
            while self._live and len(chosen) < batch_size:
                s = self._live.pop(0)
                if self._stopped[s] != 1:
                    chosen.append(s)         # finished images are dropped, as in the reference
                else:
                    self._release(s)
        idx = self._index(chosen)
        return PoolBatch(images=self.images.index_select(0, idx), states=self.states.index_select(0, idx),
                         slots=chosen, meta=[self.meta[s] for s in chosen])

    def put_back(self, slots: Sequence[int], images: torch.Tensor, states: torch.Tensor,
                 states_host: Optional[Sequence[Sequence[float]]] = None) -> int:
        """replay_memory.py:170-179: re-insert the processed records (device tensors, no host copy of
        the pixels), dropping over-long trajectories with probability 1 - over_length_keep_prob, then
        refill with fresh images.  Returns how many records were kept.

        ``states_host`` may carry the (tiny) state rows if the caller already has them on the host;
        otherwise the two fields the pool logic needs are read back (2 floats per record)."""
        if states_host is None:
            flags = states[:, [STATE_STOPPED_DIM, STATE_STEP_DIM]].detach().to("cpu").tolist()
        else:
            flags = [[row[STATE_STOPPED_DIM], row[STATE_STEP_DIM]] for row in states_host]
        keep_rows, keep_slots = [], []
        for i, s in enumerate(slots):
            stopped, step = flags[i]
            if step < self.cfg.maximum_trajectory_length or self.rng.random() < self.cfg.over_length_keep_prob:
                keep_rows.append(i)
                keep_slots.append(s)
                self._stopped[s], self._steps[s] = stopped, step
            else:
                self._release(s)
        if keep_rows:
            rows = self._index(keep_rows)
            idx = self._index(keep_slots)
            self.images.index_copy_(0, idx, images.detach().index_select(0, rows))
            self.states.index_copy_(0, idx, states.detach().index_select(0, rows).to(torch.float32))
            self._live.extend(keep_slots)
        self.fill()
        self.rng.shuffle(self._live)
        return len(keep_rows)

    def discard(self, slots: Sequence[int]) -> None:
        """train.py:374-376: a bad batch (NaN / Inf / too dark / too bright) is not stored; its slots
        are refilled with fresh images."""
        for s in slots:
            self._release(s)
        self.fill()

    def _release(self, s: int) -> None:
        self.meta[s] = None
        self._free.append(s)

    def average_trajectory(self) -> float:
        """replay_memory.py:223-230 (debug statistic)."""
        return sum(self._steps[s] for s in self._live) / max(1, len(self._live))
