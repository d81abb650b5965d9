"""Autograd-aware entry points over the C ABI (``include/aisp_b200.h``).

``apply_ops`` is the one differentiable primitive: "apply op[b] with parameters P[b] to image b",
for a homogeneous batch (one filter class, what ``Filter.process/forward`` needs) or a heterogeneous
one (the policy-selected filter per sample, ``agent.py:103-116,154``).  ``chain_forward`` is the
fused multi-step forward used for fixed chains and saved-pipeline replay.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import PSTRIDE, MAX_STEPS

# op codes == enum aisp_op
OP_EXPOSURE, OP_GAMMA, OP_CCM, OP_SHARPEN, OP_NLM, OP_TONE, OP_CONTRAST, OP_SATPLUS, OP_WNB, OP_WB, \
    OP_USM, OP_COLOR, OP_SHARPEN_V2 = range(13)
NUM_PARAMS = (1, 1, 9, 1, 1, 8, 1, 1, 1, 3, 2, 24, 1)
POINTWISE = frozenset({OP_EXPOSURE, OP_GAMMA, OP_CCM, OP_TONE, OP_CONTRAST, OP_SATPLUS, OP_WNB, OP_WB, OP_COLOR})
SHARPEN = frozenset({OP_SHARPEN, OP_SHARPEN_V2, OP_USM})

FAMILY_POINTWISE, FAMILY_SHARPEN, FAMILY_NLM, FAMILY_MIXED = "pointwise", "sharpen", "nlm", "mixed"


def family_of(op: int) -> str:
    if op in POINTWISE:
        return FAMILY_POINTWISE
    if op in SHARPEN:
        return FAMILY_SHARPEN
    if op == OP_NLM:
        return FAMILY_NLM
    raise ValueError(f"unknown op {op}")


def pack_params(param: torch.Tensor, n: int) -> torch.Tensor:
    """Reference-layout parameter tensor (e.g. Tone ``[B,8,1,1,1]``) -> packed ``[B,PSTRIDE]`` row."""
    flat = param.reshape(param.shape[0], n)
    if flat.dtype != torch.float32:
        raise _lib.AispError(f"filter parameters must be float32, got {flat.dtype}")
    return torch.nn.functional.pad(flat, (0, PSTRIDE - n))


def _ops_tensor(ops, B: int, device) -> torch.Tensor:
    if isinstance(ops, int):
        return torch.full((B,), ops, dtype=torch.int32, device=device)
    if ops.dtype != torch.int32:
        ops = ops.to(torch.int32)
    if not ops.is_cuda:
        raise _lib.AispError("ops must live on the GPU (no host round trip on the hot path)")
    return ops.contiguous()


class _ApplyOps(torch.autograd.Function):
    """out[b] = [clip](process_{ops[b]}(img[b], P[b]))  with analytic backward in CUDA.

    Optionally, in the same launch set (``aisp_sequence_fwd``): the high-resolution twin ``hr`` gets the same
    op and parameters (no gradient: the reference only does this in evaluation, agent.py:155-157), and the
    ``down_hw`` block means of ``out`` are produced from the kernels' store path (differentiable: the
    gradient that reaches them -- the critic pools the retouched image, value.py:63 -- is folded into the
    upstream gradient of ``out``)."""

    @staticmethod
    def forward(ctx, img, P, ops, clip: bool, family: str, hr, down_hw):
        _lib.require_image(img, "img")
        B, _, H, W = img.shape
        if P.shape != (B, PSTRIDE) or not P.is_cuda or P.dtype != torch.float32:
            raise _lib.AispError(f"packed params must be CUDA float32 [B,{PSTRIDE}], got {tuple(P.shape)} {P.dtype}")
        P = P.contiguous()
        L = _lib.lib()
        st = _lib.stream_ptr(img.device)
        out = torch.empty_like(img)
        want_pgrad, want_igrad = ctx.needs_input_grad[1], ctx.needs_input_grad[0]
        has_nlm = family in (FAMILY_NLM, FAMILY_MIXED)
        # NLM stashes: d out/d h (parameter gradient as one dot product) and the weight sums
        # (needed by the image-gradient kernel); rows of non-NLM samples are never touched
        stash = torch.empty_like(img) if (has_nlm and want_pgrad) else None
        wsum = torch.empty((B, 1, H, W), dtype=img.dtype, device=img.device) if (has_nlm and want_igrad) else None
        hr_out = down = None
        if hr is not None:
            _lib.require_image(hr, "high_res")
            if hr.shape[0] != B:
                raise _lib.AispError("high_res must have the batch size of img")
            hr_out = torch.empty_like(hr)
        if down_hw is not None:
            oh, ow = int(down_hw[0]), int(down_hw[1])
            if H % oh or W % ow:
                raise _lib.AispError(f"block means need evenly dividing sizes, got {H}x{W} -> {oh}x{ow}")
            down = torch.empty((B, 3, oh, ow), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            if hr is not None or down is not None:
                rc = L.aisp_sequence_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), None, B, H, W, 1,
                                         int(clip), _lib.ptr(hr), _lib.ptr(hr_out),
                                         hr.shape[2] if hr is not None else 0, hr.shape[3] if hr is not None else 0,
                                         _lib.ptr(down), down.shape[2] if down is not None else 0,
                                         down.shape[3] if down is not None else 0, _lib.ptr(stash), _lib.ptr(wsum), st)
            elif family == FAMILY_POINTWISE:
                rc = L.aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), None,
                                          B, H, W, 1, int(clip), st)
            elif family == FAMILY_SHARPEN:
                rc = L.aisp_sharpen_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W, st)
            elif family == FAMILY_NLM:
                rc = L.aisp_nlm_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W,
                                    _lib.ptr(stash), _lib.ptr(wsum), st)
            else:
                rc = L.aisp_select_apply_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                             B, H, W, int(clip), _lib.ptr(stash), _lib.ptr(wsum), st)
        _lib.check(rc, f"aisp {family} forward")
        ctx.save_for_backward(img, P, ops, stash, wsum, out if wsum is not None else None)
        ctx.clip = bool(clip)
        ctx.family = FAMILY_MIXED if (hr is not None or down is not None) else family
        ctx.set_materialize_grads(False)
        if hr_out is not None:
            ctx.mark_non_differentiable(hr_out)
        return out, hr_out, down

    @staticmethod
    def backward(ctx, g, _g_hr, g_down):
        img, P, ops, stash, wsum, out = ctx.saved_tensors
        B, _, H, W = img.shape
        need_img, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_img or need_p) or (g is None and g_down is None):
            return None, None, None, None, None, None, None
        if g_down is not None:
            g_down = g_down.contiguous()
            bh, bw = H // g_down.shape[2], W // g_down.shape[3]
            pow2 = all(v > 0 and (v & (v - 1)) == 0 for v in (W, bh, bw)) and (bw >= 4 or W % 4 != 0)
            if pow2 and not need_img and ctx.family == FAMILY_MIXED and g_down.dtype == torch.float32:
                # the pooled gradient is added inside the backward kernels' loads (aisp_select_apply_bwd_pooled)
                if g is None:
                    g = torch.zeros_like(img)
                g = g.contiguous()
                gP = torch.zeros_like(P)
                sc = _lib.scratch(B, H, W, img.device)
                with torch.cuda.device(img.device):
                    rc = _lib.lib().aisp_select_apply_bwd_pooled(
                        img.data_ptr(), None, g.data_ptr(), g_down.data_ptr(), g_down.shape[2], g_down.shape[3],
                        P.data_ptr(), ops.data_ptr(), B, H, W, int(ctx.clip), _lib.ptr(stash), None, gP.data_ptr(), None,
                        None, sc.data_ptr(), sc.numel(), _lib.stream_ptr(img.device))
                _lib.check(rc, "aisp_select_apply_bwd_pooled")
                return None, (gP if need_p else None), None, None, None, None, None
            # otherwise: d(block mean)/d(pixel) = 1 / (bh * bw), spread the pooled gradient over its block
            up = (g_down * (1.0 / (bh * bw))).repeat_interleave(bh, dim=2).repeat_interleave(bw, dim=3)
            g = up if g is None else g + up
        g = g.contiguous()
        if g.dtype != torch.float32:
            raise _lib.AispError("grad_out must be float32")
        L = _lib.lib()
        st = _lib.stream_ptr(img.device)
        family = ctx.family
        gP = torch.zeros_like(P)
        gimg = torch.empty_like(img) if need_img else None
        gy = None
        if need_img and family in (FAMILY_SHARPEN, FAMILY_MIXED):
            gy = torch.empty_like(img)
        sc = _lib.scratch(B, H, W, img.device)
        with torch.cuda.device(img.device):
            if family == FAMILY_POINTWISE:
                rc = L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W,
                                          int(ctx.clip), gP.data_ptr(), _lib.ptr(gimg), sc.data_ptr(), sc.numel(), st)
            elif family == FAMILY_SHARPEN:
                rc = L.aisp_sharpen_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W,
                                        gP.data_ptr(), _lib.ptr(gimg), _lib.ptr(gy), sc.data_ptr(), sc.numel(), st)
            elif family == FAMILY_NLM:
                rc = 0
                if need_p and stash is not None:
                    rc = L.aisp_nlm_bwd(g.data_ptr(), stash.data_ptr(), ops.data_ptr(), B, H, W, gP.data_ptr(),
                                        sc.data_ptr(), sc.numel(), st)
                if rc == 0 and need_img:
                    rc = L.aisp_nlm_bwd_img(img.data_ptr(), out.data_ptr(), wsum.data_ptr(), g.data_ptr(), P.data_ptr(),
                                            ops.data_ptr(), B, H, W, gimg.data_ptr(), st)
            else:
                rc = L.aisp_select_apply_bwd(img.data_ptr(), _lib.ptr(out), g.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                             B, H, W, int(ctx.clip), _lib.ptr(stash), _lib.ptr(wsum), gP.data_ptr(),
                                             _lib.ptr(gimg), _lib.ptr(gy), sc.data_ptr(), sc.numel(), st)
        _lib.check(rc, f"aisp {family} backward")
        return gimg, (gP if need_p else None), None, None, None, None, None


def apply_ops(img: torch.Tensor, P: torch.Tensor, ops, clip: bool, family: Optional[str] = None,
              high_res: Optional[torch.Tensor] = None, down_hw=None):
    """Apply ``ops[b]`` (int or int32 CUDA tensor ``[B]``) with packed parameters ``P[b]``.

    ``clip=True`` reproduces ``Filter.forward`` (isp/filters.py:115,125), ``clip=False``
    ``Filter.process`` / ``run`` (:138).  ``family`` narrows the launch to one kernel family when the
    caller knows the batch is homogeneous; ``None`` means heterogeneous ("mixed": three launches).

    ``high_res`` ``[B,3,H2,W2]``: the full-size twin gets the same op and parameters in the same launch set
    (isp/filters.py:116-122, agent.py:155-157).  ``down_hw`` ``(oh, ow)``: also return the block means of the
    output (== ``nn.AdaptiveAvgPool2d`` for evenly dividing sizes), emitted from the kernels' store path.
    With either, the result is ``(out, high_res_out | None, down | None)``; otherwise just ``out``.
    """
    if isinstance(ops, int) and family is None:
        family = family_of(ops)
    ops_t = _ops_tensor(ops, img.shape[0], img.device)
    out, hr_out, down = _ApplyOps.apply(img, P, ops_t, clip, family or FAMILY_MIXED, high_res, down_hw)
    if high_res is None and down_hw is None:
        return out
    return out, hr_out, down


def apply_filter(img: torch.Tensor, param: torch.Tensor, op: int, clip: bool) -> torch.Tensor:
    """One filter class on the whole batch; ``param`` in the reference's own layout."""
    return apply_ops(img, pack_params(param, NUM_PARAMS[op]), op, clip, family_of(op))


class _ApplyBank(torch.autograd.Function):
    """stack[b, f] = [clip](process_{ops[f]}(img[b], P[b, f])) for F filters on the same batch."""

    @staticmethod
    def forward(ctx, img, P, ops, clip: bool):
        _lib.require_image(img, "img")
        B, _, H, W = img.shape
        F = len(ops)
        has_nlm = OP_NLM in ops
        ops_c = (ctypes.c_int32 * F)(*ops)          # host array: the bank's op list never lives on the device
        if ctx.needs_input_grad[0]:
            raise _lib.AispError("apply_bank differentiates w.r.t. the filter parameters only; "
                                 "detach the image (train.py:255) or use apply_ops per filter")
        out = torch.empty((B, F, 3, H, W), dtype=torch.float32, device=img.device)
        stash = torch.empty_like(img) if (has_nlm and ctx.needs_input_grad[1]) else None
        with torch.cuda.device(img.device):
            rc = _lib.lib().aisp_bank_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops_c, B, F, H, W,
                                          int(clip), _lib.ptr(stash), _lib.stream_ptr(img.device))
        _lib.check(rc, "aisp_bank_fwd")
        ctx.save_for_backward(img, P, stash)
        ctx.clip = bool(clip)
        ctx.ops_c = ops_c
        return out

    @staticmethod
    def backward(ctx, g):
        img, P, stash = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None, None, None
        B, _, H, W = img.shape
        F = len(ctx.ops_c)
        g = g.contiguous()
        if g.dtype != torch.float32:
            raise _lib.AispError("grad_out must be float32")
        gP = torch.zeros_like(P)
        sc = _lib.scratch(B * F, H, W, img.device)
        with torch.cuda.device(img.device):
            rc = _lib.lib().aisp_bank_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ctx.ops_c, B, F, H, W,
                                          int(ctx.clip), _lib.ptr(stash), gP.data_ptr(), sc.data_ptr(), sc.numel(),
                                          _lib.stream_ptr(img.device))
        _lib.check(rc, "aisp_bank_bwd")
        return None, gP, None, None


def apply_bank(img: torch.Tensor, P: torch.Tensor, ops: Sequence[int], clip: bool = True) -> torch.Tensor:
    """All of ``ops`` (F op codes) applied to the same batch, results stacked on dim 1 -- the
    ``torch.stack([f(img, ...) for f in filters], dim=1)`` of agent.py:103-107 -- in three launches
    (one per kernel family): the image is fetched from HBM once per direction, not once per filter,
    and the values are bit-identical to F separate :func:`apply_ops` calls.

    img ``[B,3,H,W]`` (no gradient flows to it); P ``[B,F,PSTRIDE]`` packed rows (may require grad);
    returns ``[B,F,3,H,W]``.  At most one NLM entry in ``ops`` (its d/dh stash is kept compact).
    """
    ops = tuple(int(o) for o in ops)
    B, F = img.shape[0], len(ops)
    if not (1 <= F <= _lib.MAX_BANK_FILTERS) or B * F > 65535:
        raise _lib.AispError(f"bank of {F} filters on a batch of {B}: need 1 <= F <= {_lib.MAX_BANK_FILTERS} "
                             "and B*F <= 65535")
    if P.shape != (B, F, PSTRIDE) or P.dtype != torch.float32 or not P.is_cuda:
        raise _lib.AispError(f"P must be CUDA float32 [B,F,{PSTRIDE}], got {tuple(P.shape)} {P.dtype}")
    if any(not (0 <= o < len(NUM_PARAMS)) for o in ops):
        raise _lib.AispError(f"unknown op code in {ops}")
    n_nlm = sum(1 for o in ops if o == OP_NLM)
    if n_nlm > 1:
        raise _lib.AispError("a filter bank holds at most one NLM filter")
    return _ApplyBank.apply(img, P.contiguous(), ops, clip)


class _Regress(torch.autograd.Function):
    """raw fc_filter outputs of a whole bank [B,Ntot] -> packed parameter rows [B,F,PSTRIDE], one launch each
    way (``aisp_regress_fwd/bwd``): every filter's ``filter_param_regressor`` (isp/filters.py:215-708)."""

    @staticmethod
    def forward(ctx, raw, fops, offs, cfg_c, F: int):
        if not raw.is_cuda or raw.dtype != torch.float32 or raw.dim() != 2:
            raise _lib.AispError("regress: raw features must be a CUDA float32 [B,Ntot] tensor")
        raw = raw.contiguous()
        B, Ntot = raw.shape
        packed = torch.empty((B, F, PSTRIDE), dtype=torch.float32, device=raw.device)
        with torch.cuda.device(raw.device):
            rc = _lib.lib().aisp_regress_fwd(raw.data_ptr(), fops.data_ptr(), offs.data_ptr(), B, F, Ntot, cfg_c,
                                             packed.data_ptr(), _lib.stream_ptr(raw.device))
        _lib.check(rc, "aisp_regress_fwd")
        ctx.save_for_backward(raw, fops, offs)
        ctx.cfg_c, ctx.F = cfg_c, F
        return packed

    @staticmethod
    def backward(ctx, g):
        raw, fops, offs = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        B, Ntot = raw.shape
        g = g.contiguous()
        graw = torch.empty_like(raw)
        with torch.cuda.device(raw.device):
            rc = _lib.lib().aisp_regress_bwd(raw.data_ptr(), g.data_ptr(), fops.data_ptr(), offs.data_ptr(), B, ctx.F,
                                             Ntot, ctx.cfg_c, graw.data_ptr(), _lib.stream_ptr(raw.device))
        _lib.check(rc, "aisp_regress_bwd")
        return graw, None, None, None, None


def regress_ranges(cfg):
    """The 15 floats of ``aisp_regress_*``'s ``cfg_ranges``: (lo, span) pairs formed in Python floats exactly as
    the reference's ``tanh_range(l, r, initial)`` forms ``r - l`` and its ``atanh`` shift (isp/filters.py:27-34)."""
    import math
    lg = math.log(cfg.gamma_range)
    cl, ch = cfg.color_curve_range
    vals = [-cfg.exposure_range, cfg.exposure_range - (-cfg.exposure_range), -lg, lg - (-lg),
            cfg.tone_curve_range[0], cfg.tone_curve_range[1] - cfg.tone_curve_range[0],
            cl, ch - cl, math.atanh(2 * (1 - cl) / (ch - cl) - 1),
            cfg.usm_sharpen_range[0], cfg.usm_sharpen_range[1] - cfg.usm_sharpen_range[0],
            cfg.sharpen_range[0], cfg.sharpen_range[1] - cfg.sharpen_range[0],
            cfg.ccm_range[0], cfg.ccm_range[1] - cfg.ccm_range[0]]
    return (ctypes.c_float * 15)(*[float(v) for v in vals])


def regress(raw: torch.Tensor, fops: torch.Tensor, offs: torch.Tensor, cfg_c, F: int) -> torch.Tensor:
    """Differentiable: ``raw`` [B,Ntot] -> packed [B,F,PSTRIDE] (see :class:`_Regress`)."""
    return _Regress.apply(raw, fops, offs, cfg_c, F)


SELECT_SAMPLE, SELECT_ARGMAX, SELECT_FORCED = 0, 1, 2   # enum aisp_select_mode


class _SelectRows(torch.autograd.Function):
    """Selection + one-hot + parameter-row gather + state update in one launch (``aisp_select``).
    Differentiable w.r.t. ``packed_all`` only (the gather); everything else is integer bookkeeping."""

    @staticmethod
    def forward(ctx, pdf, noise, states, packed_all, op_table, mode: int, forced: int, test_steps: float,
                early_stop_c: float):
        B, F = pdf.shape
        S = states.shape[1]
        dev = pdf.device
        for name, t in (("pdf", pdf), ("states", states), ("packed_all", packed_all)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise _lib.AispError(f"{name} must be a CUDA float32 tensor")
        if packed_all.shape != (B, F, PSTRIDE) or S != 3 + F or op_table.numel() != F:
            raise _lib.AispError("select: expected pdf [B,F], states [B,3+F], packed_all [B,F,24], op_table [F]")
        pdf, states, packed_all = pdf.contiguous(), states.contiguous(), packed_all.contiguous()
        if noise is not None:
            noise = noise.reshape(B).to(torch.float32).contiguous()
        sel = torch.empty((B,), dtype=torch.int64, device=dev)
        hot = torch.empty((B, F), dtype=torch.int64, device=dev)
        ops = torch.empty((B,), dtype=torch.int32, device=dev)
        rows = torch.empty((B, PSTRIDE), dtype=torch.float32, device=dev)
        new_states = torch.empty((B, S), dtype=torch.float32, device=dev)
        pens = torch.empty((B, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().aisp_select(pdf.data_ptr(), _lib.ptr(noise), int(mode), int(forced), states.data_ptr(),
                                        packed_all.data_ptr(), op_table.data_ptr(), B, F, S, float(test_steps),
                                        float(early_stop_c), sel.data_ptr(), hot.data_ptr(), ops.data_ptr(),
                                        rows.data_ptr(), new_states.data_ptr(), pens.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "aisp_select")
        ctx.save_for_backward(sel)
        ctx.F = F
        ctx.mark_non_differentiable(sel, hot, ops, new_states, pens)
        return rows, sel, hot, ops, new_states, pens

    @staticmethod
    def backward(ctx, g_rows, *_):
        (sel,) = ctx.saved_tensors
        if not ctx.needs_input_grad[3]:
            return (None,) * 9
        B, F = sel.shape[0], ctx.F
        g_rows = g_rows.contiguous()
        g_all = torch.empty((B, F, PSTRIDE), dtype=torch.float32, device=sel.device)
        with torch.cuda.device(sel.device):
            rc = _lib.lib().aisp_select_bwd(g_rows.data_ptr(), sel.data_ptr(), B, F, g_all.data_ptr(),
                                            _lib.stream_ptr(sel.device))
        _lib.check(rc, "aisp_select_bwd")
        return None, None, None, g_all, None, None, None, None, None


def select_rows(pdf, noise, states, packed_all, op_table, mode: int, forced: int = 0, test_steps: float = 5.0,
                early_stop_c: float = 1.0):
    """Device-side selection step (agent.py:12-23,126-154,234-259) -> ``(rows [B,24], sel [B] int64,
    one_hot [B,F] int64, ops [B] int32, new_states [B,S], penalties [B,2] = (usage, early stop))``.
    ``mode``: ``SELECT_SAMPLE`` (training, needs ``noise`` [B] or [B,1]), ``SELECT_ARGMAX`` (eval) or
    ``SELECT_FORCED`` (``forced`` = filter index).  ``pdf`` is read as data (no gradient through the
    integer choice); ``rows`` carries the gradient back to ``packed_all``."""
    return _SelectRows.apply(pdf.detach(), noise, states.detach(), packed_all, op_table, mode, forced, test_steps,
                             early_stop_c)


class _ApplyChain(torch.autograd.Function):
    """Per-sample sequences of per-pixel filters, forward AND backward each fused into one pass."""

    @staticmethod
    def forward(ctx, img, P, ops, seq_len, clip_each: bool):
        _lib.require_image(img, "img")
        B, _, H, W = img.shape
        S = ops.shape[1]
        out = torch.empty_like(img)
        with torch.cuda.device(img.device):
            rc = _lib.lib().aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                               _lib.ptr(seq_len), B, H, W, S, int(clip_each) | _lib.SEQ_STRICT,
                                               _lib.stream_ptr(img.device))
        _lib.check(rc, "aisp_pointwise_fwd")
        ctx.save_for_backward(img, P, ops, seq_len)
        ctx.clip_each = bool(clip_each)
        return out

    @staticmethod
    def backward(ctx, g):
        img, P, ops, seq_len = ctx.saved_tensors
        need_img, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_img or need_p):
            return None, None, None, None, None
        B, _, H, W = img.shape
        S = ops.shape[1]
        g = g.contiguous()
        gP = torch.empty_like(P)
        gimg = torch.empty_like(img) if need_img else None
        sc = _lib.scratch(B, H, W, img.device)
        with torch.cuda.device(img.device):
            rc = _lib.lib().aisp_pointwise_chain_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                                     _lib.ptr(seq_len), B, H, W, S,
                                                     int(ctx.clip_each) | _lib.SEQ_STRICT, gP.data_ptr(),
                                                     _lib.ptr(gimg), sc.data_ptr(), sc.numel(),
                                                     _lib.stream_ptr(img.device))
        _lib.check(rc, "aisp_pointwise_chain_bwd")
        return gimg, (gP if need_p else None), None, None, None


def apply_chain(img: torch.Tensor, P: torch.Tensor, ops: torch.Tensor, seq_len: Optional[torch.Tensor] = None,
                clip_each: bool = True) -> torch.Tensor:
    """Differentiable fused sequence: ``ops[b, :seq_len[b]]`` (per-pixel filters only) applied to image b
    with parameter rows ``P[b, k]``; forward in one pass over HBM, backward in one pass.

    img ``[B,3,H,W]``; P ``[B,S,PSTRIDE]`` (may require grad); ops int32 ``[B,S]`` with ``S <= MAX_CHAIN_BWD``
    (6) when gradients are needed.  ``clip_each`` as in :func:`chain_forward`.  Every per-pixel op is
    differentiated, ColorFilter included.  The ops live on the device and are not read back: a sample
    whose sequence holds a stencil op (Shr / USM / NLM -- use :func:`run_pipeline` or the replay plan
    for those) or an unknown code gets an all-NaN output and NaN gradients, never stale memory.
    """
    B = img.shape[0]
    S = ops.shape[1]
    if P.shape != (B, S, PSTRIDE) or P.dtype != torch.float32 or not P.is_cuda:
        raise _lib.AispError(f"P must be CUDA float32 [B,S,{PSTRIDE}]")
    needs_grad = torch.is_grad_enabled() and (img.requires_grad or P.requires_grad)
    if needs_grad and S > _lib.MAX_CHAIN_BWD:
        raise _lib.AispError(f"a fused backward sequence holds at most {_lib.MAX_CHAIN_BWD} steps (got {S}); "
                             "split the chain and chain the calls")
    if not needs_grad and not (1 <= S <= MAX_STEPS):
        raise _lib.AispError(f"sequence length {S} outside 1..{MAX_STEPS}")
    ops = _ops_tensor(ops, B, img.device)
    if seq_len is not None:
        seq_len = _ops_tensor(seq_len, B, img.device)
    return _ApplyChain.apply(img, P.contiguous(), ops, seq_len, clip_each)


@torch.no_grad()
def block_mean(img: torch.Tensor, out_hw=(64, 64)) -> torch.Tensor:
    """``nn.AdaptiveAvgPool2d(out_hw)`` for evenly dividing sizes, in one streaming read (no autograd:
    the reference only pools images that carry no gradient, agent.py:97 / train.py:255)."""
    _lib.require_image(img, "img")
    B, _, H, W = img.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    down = torch.empty((B, 3, oh, ow), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        rc = _lib.lib().aisp_block_mean(img.data_ptr(), down.data_ptr(), B, H, W, oh, ow, _lib.stream_ptr(img.device))
    _lib.check(rc, "aisp_block_mean")
    return down


@torch.no_grad()
def value_stats(down: torch.Tensor) -> torch.Tensor:
    """The critic's statistics of the pooled image (value.py:64-75) -> ``[B,3]`` = (mean luminance, unbiased
    luminance variance, mean saturation), one launch instead of ~20 (no autograd: use
    ``value.value_statistics`` where a gradient must flow)."""
    _lib.require_image(down, "down")
    B, _, h, w = down.shape
    stats = torch.empty((B, 3), dtype=torch.float32, device=down.device)
    with torch.cuda.device(down.device):
        rc = _lib.lib().aisp_value_stats(down.data_ptr(), B, h, w, stats.data_ptr(), _lib.stream_ptr(down.device))
    _lib.check(rc, "aisp_value_stats")
    return stats


def image_stats(down: torch.Tensor):
    """Per-image mean and finiteness from the block-mean image (train.py:288-290,374): the mean of
    equal-size block means is the image mean; a block mean is finite iff its whole block is."""
    return down.mean(dim=(1, 2, 3)), torch.isfinite(down).flatten(1).all(dim=1)


@torch.no_grad()
def chain_forward(img: torch.Tensor, P: torch.Tensor, ops: torch.Tensor, seq_len: Optional[torch.Tensor] = None,
                  clip_each: bool = True, strict: bool = True) -> torch.Tensor:
    """Per-sample sequences of per-pixel filters fused into ONE pass over HBM (forward only).

    img ``[B,3,H,W]``; P ``[B,S,PSTRIDE]``; ops int32 ``[B,S]``; seq_len int32 ``[B]`` or None.
    Stencil ops are not allowed inside a fused sequence (use ``run_pipeline`` for mixed sequences):
    with ``strict`` (default) such a sample's output is all NaN; ``strict=False`` gives the C ABI's
    select-apply behaviour (samples led by a stencil op are left untouched for the stencil launches of
    the same phase -- what :mod:`replay` relies on -- and a later stencil op ends the sequence).
    """
    _lib.require_image(img, "img")
    B, _, H, W = img.shape
    S = ops.shape[1]
    if not (1 <= S <= MAX_STEPS):
        raise _lib.AispError(f"sequence length {S} outside 1..{MAX_STEPS}")
    if P.shape != (B, S, PSTRIDE) or P.dtype != torch.float32 or not P.is_cuda:
        raise _lib.AispError(f"P must be CUDA float32 [B,S,{PSTRIDE}]")
    ops = _ops_tensor(ops, B, img.device)
    if seq_len is not None:
        seq_len = _ops_tensor(seq_len, B, img.device)
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        rc = _lib.lib().aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), P.contiguous().data_ptr(), ops.data_ptr(),
                                           _lib.ptr(seq_len), B, H, W, S,
                                           int(clip_each) | (_lib.SEQ_STRICT if strict else 0),
                                           _lib.stream_ptr(img.device))
    _lib.check(rc, "aisp_pointwise_fwd")
    return out


@torch.no_grad()
def sequence_forward(img: torch.Tensor, P: torch.Tensor, ops: torch.Tensor, seq_len: Optional[torch.Tensor] = None,
                     clip_each: bool = True, high_res: Optional[torch.Tensor] = None, down_hw=None):
    """Per-sample op sequences with AT MOST ONE stencil step each (Shr / ShrV2 / USM / NLM, anywhere in the
    sequence) in one launch set -- ``aisp_sequence_fwd``: per-pixel steps before the stencil step run on the
    staged tile, steps after it on the outputs in registers, so a whole sequence costs one pass over HBM.
    A second stencil step ends a sample's sequence (``replay.plan_pipeline`` splits such pipelines).

    img ``[B,3,H,W]``; P ``[B,S,PSTRIDE]``; ops int32 ``[B,S]``; seq_len int32 ``[B]`` or None.
    Returns ``(out, high_res_out | None, down | None)``."""
    _lib.require_image(img, "img")
    B, _, H, W = img.shape
    S = ops.shape[1]
    if not (1 <= S <= MAX_STEPS):
        raise _lib.AispError(f"sequence length {S} outside 1..{MAX_STEPS}")
    if P.shape != (B, S, PSTRIDE) or P.dtype != torch.float32 or not P.is_cuda:
        raise _lib.AispError(f"P must be CUDA float32 [B,S,{PSTRIDE}]")
    ops = _ops_tensor(ops, B, img.device)
    if seq_len is not None:
        seq_len = _ops_tensor(seq_len, B, img.device)
    out = torch.empty_like(img)
    hr_out = down = None
    if high_res is not None:
        _lib.require_image(high_res, "high_res")
        hr_out = torch.empty_like(high_res)
    if down_hw is not None:
        down = torch.empty((B, 3, int(down_hw[0]), int(down_hw[1])), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        rc = _lib.lib().aisp_sequence_fwd(
            img.data_ptr(), out.data_ptr(), P.contiguous().data_ptr(), ops.data_ptr(), _lib.ptr(seq_len), B, H, W, S,
            int(clip_each), _lib.ptr(high_res), _lib.ptr(hr_out), high_res.shape[2] if high_res is not None else 0,
            high_res.shape[3] if high_res is not None else 0, _lib.ptr(down), down.shape[2] if down is not None else 0,
            down.shape[3] if down is not None else 0, None, None, _lib.stream_ptr(img.device))
    _lib.check(rc, "aisp_sequence_fwd")
    return out, hr_out, down


@torch.no_grad()
def run_pipeline(img: torch.Tensor, steps: Sequence[Sequence[int]], params: Sequence[Sequence[torch.Tensor]],
                 clip_each: bool = True) -> torch.Tensor:
    """Replay known per-sample pipelines (``param_results/*.json`` of yolov3/val_adaptiveisp.py:301-327).

    ``steps[b]`` is the list of op codes of sample b, ``params[b][k]`` the flat parameter tensor of its
    k-th step.  Consecutive per-pixel steps are fused into one pass; a stencil step forces a pass
    boundary.  Samples are advanced phase by phase: phase p runs every sample's p-th fused segment in
    one heterogeneous launch (samples that already finished are carried through unchanged).
    """
    from . import replay   # planning + execution live there; this is the one-shot convenience form
    return replay.execute_plan(img, replay.plan_pipeline(steps, params, img.device), clip_each)
