"""Differentiable replay: backward through a planned batch of pipelines.

``replay.execute_plan`` is the forward-only fast path (one sequence launch set per phase).  Here the same
plan is run so that autograd can go back through it -- w.r.t. the input batch and w.r.t. every step's
parameters -- e.g. to fine-tune the parameters of saved pipelines (``param_results/*.json``) against a
loss on the final image, or to differentiate the fixed chain of isp/filters.py:753-815.

A phase's sequences hold per-pixel steps around at most one stencil step.  Differentiably, a phase is

    fused per-pixel prefix  ->  the stencil step (on the samples that have one)  ->  fused per-pixel suffix

where prefix and suffix are ``functional.apply_chain`` (forward one pass, backward one pass: the fused
sequence backward differentiates up to 6 stages per launch, longer runs are cut in two) and the stencil
step is ``functional.apply_ops`` on the gathered sub-batch.  The plan is made on the host, so which
samples have a stencil step in which phase is known when the plan is built: no device read-back.

    plan = replay.plan_pipeline(steps, params, device)
    P = [ph.params.clone().requires_grad_(True) for ph in plan.phases]     # [B,S,PSTRIDE] per phase
    out = apply_plan(img, plan, P)
    loss(out).backward()                                                   # P[i].grad[b, k, :n]
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from . import functional as AF
from ._lib import MAX_CHAIN_BWD, PSTRIDE
from .replay import PipelinePlan


@dataclass
class _PhaseIndex:
    pre_ops: Optional[torch.Tensor]      # int32 [B,Sp]  (None: no sample has a prefix)
    pre_len: Optional[torch.Tensor]      # int32 [B]
    st_idx: Optional[torch.Tensor]       # int64 [n]   samples with a stencil step
    st_pos: Optional[torch.Tensor]       # int64 [n]   its position in the phase's sequence
    st_ops: Optional[torch.Tensor]       # int32 [n]
    suf_ops: Optional[torch.Tensor]      # int32 [B,Ss]
    suf_len: Optional[torch.Tensor]      # int32 [B]
    suf_gather: Optional[torch.Tensor]   # int64 [B,Ss]  column of the phase's params for suffix step j


def _index_phase(ph) -> _PhaseIndex:
    """Host-side split of one phase into prefix / stencil / suffix index tensors (once per plan)."""
    ops = ph.ops.cpu()
    lens = ph.seq_len.cpu().tolist()
    B, S = ops.shape
    dev = ph.ops.device
    pre_len, st_idx, st_pos, st_ops, suf_len = [0] * B, [], [], [], [0] * B
    for b in range(B):
        n, pos = lens[b], -1
        for k in range(n):
            if int(ops[b, k]) not in AF.POINTWISE:
                pos = k
                break
        if pos < 0:
            pre_len[b] = n
        else:
            pre_len[b] = pos
            st_idx.append(b)
            st_pos.append(pos)
            st_ops.append(int(ops[b, pos]))
            suf_len[b] = n - pos - 1
    idx = _PhaseIndex(None, None, None, None, None, None, None, None)
    Sp, Ss = max(pre_len, default=0), max(suf_len, default=0)
    if Sp:
        idx.pre_ops = ops[:, :Sp].contiguous().to(dev)
        idx.pre_len = torch.tensor(pre_len, dtype=torch.int32, device=dev)
    if st_idx:
        idx.st_idx = torch.tensor(st_idx, dtype=torch.long, device=dev)
        idx.st_pos = torch.tensor(st_pos, dtype=torch.long, device=dev)
        idx.st_ops = torch.tensor(st_ops, dtype=torch.int32, device=dev)
    if Ss:
        gather = torch.zeros((B, Ss), dtype=torch.long)
        sops = torch.zeros((B, Ss), dtype=torch.int32)
        for b, pos in zip(st_idx, st_pos):
            for j in range(suf_len[b]):
                gather[b, j] = pos + 1 + j
                sops[b, j] = ops[b, pos + 1 + j]
        idx.suf_ops = sops.to(dev)
        idx.suf_len = torch.tensor(suf_len, dtype=torch.int32, device=dev)
        idx.suf_gather = gather.to(dev)
    return idx


def _chain(x, P, ops, lens, clip_each):
    """apply_chain in pieces of at most MAX_CHAIN_BWD steps (the fused backward's limit)."""
    S = ops.shape[1]
    for lo in range(0, S, MAX_CHAIN_BWD):
        hi = min(S, lo + MAX_CHAIN_BWD)
        piece_len = torch.clamp(lens - lo, min=0, max=hi - lo).to(torch.int32)
        x = AF.apply_chain(x, P[:, lo:hi].contiguous(), ops[:, lo:hi].contiguous(), piece_len, clip_each=clip_each)
    return x


def apply_plan(img: torch.Tensor, plan: PipelinePlan, params: Optional[Sequence[torch.Tensor]] = None,
               clip_each: bool = True) -> torch.Tensor:
    """Differentiable ``execute_plan``: gradients flow to ``img`` (if it requires grad) and to
    ``params[i]`` (``[B,S_i,PSTRIDE]`` per phase, default ``plan.phases[i].params``).  Values agree with
    ``execute_plan`` to rounding (the per-pixel steps use the same arithmetic; only the launch structure
    differs)."""
    _lib.require_image(img, "img")
    if img.shape[0] != plan.batch:
        raise _lib.AispError(f"plan was made for batch {plan.batch}, got {img.shape[0]}")
    if params is None:
        params = [ph.params for ph in plan.phases]
    if len(params) != len(plan.phases):
        raise _lib.AispError("one parameter tensor per phase is required")
    index: List[_PhaseIndex] = getattr(plan, "_grad_index", None)
    if index is None:
        index = [_index_phase(ph) for ph in plan.phases]
        plan._grad_index = index
    x = img
    for ph, P, ix in zip(plan.phases, params, index):
        if P.shape != ph.params.shape:
            raise _lib.AispError(f"phase parameters must be {tuple(ph.params.shape)}, got {tuple(P.shape)}")
        if ix.pre_ops is not None:
            x = _chain(x, P[:, :ix.pre_ops.shape[1]], ix.pre_ops, ix.pre_len, clip_each)
        if ix.st_idx is not None:
            rows = P[ix.st_idx, ix.st_pos]                                   # [n,PSTRIDE], differentiable gather
            sub = AF.apply_ops(x.index_select(0, ix.st_idx), rows, ix.st_ops, clip=clip_each)
            x = x.index_copy(0, ix.st_idx, sub)
        if ix.suf_ops is not None:
            Ps = torch.gather(P, 1, ix.suf_gather[:, :, None].expand(-1, -1, PSTRIDE))
            x = _chain(x, Ps, ix.suf_ops, ix.suf_len, clip_each)
    return x
