"""Replay of known per-image filter pipelines in as few passes over HBM as possible.

A live rollout is sequentially dependent (each step's parameters are predicted from the previous
step's output, yolov3/val_adaptiveisp.py:291-309), but once the (filter id, parameters) sequence of
an image is known -- the ``param_results/<img>.json`` and ``records.txt`` files that the reference's
evaluation script writes (:269-270, :301-327), or a ``--pipeline`` forced sequence -- whole sequences
can be fused: consecutive per-pixel filters of a sample become ONE pass (``aisp_pointwise_fwd`` with a
per-sample op sequence), a stencil filter (3x3 sharpen, USM, NLM) forces a pass boundary, and samples
with different sequences share the same launches (heterogeneous per-sample dispatch).

    plan = plan_pipeline(steps, params, device)     # once: host-side segmentation + one upload
    out  = execute_plan(img, plan, clip_each=True)  # per batch: no host work, no sync
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import functional as AF
from ._lib import MAX_STEPS, PSTRIDE


@dataclass
class Phase:
    ops: torch.Tensor                 # int32 [B,S]
    seq_len: torch.Tensor             # int32 [B]   (0 = sample idle in this phase: copied through)
    params: torch.Tensor              # fp32  [B,S,PSTRIDE]
    has_sharpen: bool = False         # informational: which stencil families appear in the phase
    has_nlm: bool = False


@dataclass
class PipelinePlan:
    batch: int
    phases: List[Phase] = field(default_factory=list)
    launches: int = 0                 # kernel launches with work (idle family launches exit at once)


def segment(ops: Sequence[int]) -> List[List[int]]:
    """Indices of one sample's steps grouped into fused segments: a segment holds per-pixel steps and AT
    MOST ONE stencil step (anywhere in it) and is at most MAX_STEPS long -- what one
    ``aisp_sequence_fwd`` call runs in a single pass over HBM.  A second stencil step opens a new segment
    (its halo would need the first stencil's output at neighbouring pixels)."""
    out, cur, has_stencil = [], [], False
    for k, op in enumerate(ops):
        stencil = op not in AF.POINTWISE
        if cur and (len(cur) == MAX_STEPS or (stencil and has_stencil)):
            out.append(cur)
            cur, has_stencil = [], False
        cur.append(k)
        has_stencil = has_stencil or stencil
    if cur:
        out.append(cur)
    return out


def plan_pipeline(steps: Sequence[Sequence[int]], params: Sequence[Sequence[torch.Tensor]], device) -> PipelinePlan:
    """``steps[b]``: op codes (``aisp_op``) of sample b in order; ``params[b][k]``: the parameter tensor of
    its k-th step (any shape with the filter's number of elements)."""
    B = len(steps)
    segs = [segment(s) for s in steps]
    nphase = max((len(s) for s in segs), default=0)
    plan = PipelinePlan(batch=B)
    for p in range(nphase):
        S = max(len(s[p]) if p < len(s) else 0 for s in segs)
        ops_h = torch.zeros((B, S), dtype=torch.int32)
        len_h = torch.zeros((B,), dtype=torch.int32)
        P_h = torch.zeros((B, S, PSTRIDE), dtype=torch.float32)
        fam = set()
        for b in range(B):
            if p >= len(segs[b]):
                continue
            f = AF.FAMILY_POINTWISE
            for j, k in enumerate(segs[b][p]):
                op = int(steps[b][k])
                ops_h[b, j] = op
                P_h[b, j, :AF.NUM_PARAMS[op]] = params[b][k].detach().reshape(-1).float().cpu()
                if op not in AF.POINTWISE:
                    f = AF.family_of(op)
            len_h[b] = len(segs[b][p])
            fam.add(f)
        ph = Phase(ops=ops_h.to(device), seq_len=len_h.to(device), params=P_h.to(device))
        ph.has_sharpen = AF.FAMILY_SHARPEN in fam
        ph.has_nlm = AF.FAMILY_NLM in fam
        plan.phases.append(ph)
        plan.launches += max(1, len(fam))
    return plan


@torch.no_grad()
def execute_plan(img: torch.Tensor, plan: PipelinePlan, clip_each: bool = True,
                 high_res: Optional[torch.Tensor] = None):
    """Run a planned batch of pipelines: one sequence launch set per phase (``aisp_sequence_fwd``: the
    per-pixel, sharpen and NLM kernels back to back, each sample served by exactly one of them; samples
    that already finished are carried through unchanged).  With ``high_res`` the same pipelines are
    applied to the full-size twins in the same launches and ``(out, high_res_out)`` is returned."""
    _lib.require_image(img, "img")
    if img.shape[0] != plan.batch:
        raise _lib.AispError(f"plan was made for batch {plan.batch}, got {img.shape[0]}")
    if not plan.phases:
        return img.clone() if high_res is None else (img.clone(), high_res.clone())
    x, hr = img, high_res
    for ph in plan.phases:
        x, hr, _ = AF.sequence_forward(x, ph.params, ph.ops, ph.seq_len, clip_each, high_res=hr)
    return x if high_res is None else (x, hr)


# ------------------------------------------------------------------------------------------------
# the reference's on-disk formats
# ------------------------------------------------------------------------------------------------
def parse_param_results(obj, filters) -> Tuple[List[int], List[torch.Tensor]]:
    """One ``param_results/<img>.json`` (path, JSON text or dict) -> (op codes, parameter tensors).

    Format (yolov3/val_adaptiveisp.py:283-327): ``{"pipeline": [filter ids...], "<short_name>":
    nested-list parameters, ...}`` with ids indexing ``cfg.filters``.  ``filters`` is that list (classes
    or instances of the drop-in filters).  The file keeps ONE parameter entry per filter name, so a
    filter used twice replays with its last recorded parameters -- as the file itself implies."""
    if isinstance(obj, str):
        obj = json.loads(obj) if obj.lstrip().startswith("{") else json.load(open(obj))
    names = [(f.short_name if not isinstance(f, type) else _short_name_of(f)) for f in filters]
    ops, plist = [], []
    for fid in obj["pipeline"]:
        fid = int(fid)
        flt = filters[fid]
        ops.append(int(flt.OP))
        plist.append(torch.tensor(obj[names[fid]], dtype=torch.float32).reshape(-1))
    return ops, plist


def _short_name_of(cls) -> str:
    table = {"ExposureFilter": "E", "GammaFilter": "G", "CCMFilter": "CCM", "SharpenFilter": "Shr",
             "DenoiseFilter": "NLM", "ToneFilter": "T", "ContrastFilter": "Ct", "SaturationPlusFilter": "S+",
             "WNBFilter": "BW", "ImprovedWhiteBalanceFilter": "W", "SharpenUSMFilter": "USM", "ColorFilter": "C",
             "SharpenFilterV2": "Shr", "ToneFilterV2": "T"}
    return table[cls.__name__]


def parse_records(text: str) -> Tuple[List[str], Dict[str, List[int]]]:
    """``records.txt`` (yolov3/val_adaptiveisp.py:269-270, :321-322): header line = filter short names,
    then ``<file>,<id0>,...,<idN>`` with -1 for steps that were not run."""
    lines = [l.strip() for l in text.splitlines() if l.strip()]
    header = lines[0].split(",")
    rec = {}
    for l in lines[1:]:
        parts = l.split(",")
        rec[parts[0]] = [int(v) for v in parts[1:] if int(v) >= 0]
    return header, rec
