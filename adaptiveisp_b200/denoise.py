"""Drop-in for the reference's ``isp/denoise.py``: the non-local-means MODULES themselves (SURVEY §8a rows
A11 / A12), backed by the same sm_100a kernel as ``DenoiseFilter`` (``csrc/nlm.cu``).

``NonLocalMeansGray`` (isp/denoise.py:93-119) is what ``DenoiseFilter`` instantiates (isp/filters.py:577);
``NonLocalMeans`` (:68-90, per-channel distances and weights) is its commented-out alternative (:576).
Both are built for the one window configuration the reference uses, ``search_window_size=11, patch_size=5``
(the kernel's tile geometry is that of an 11x11 search and a 5x5 patch); other sizes raise.  Gradients:
w.r.t. ``h`` (closed form accumulated in the forward pass, as for the filter); an image that requires grad
raises -- use ``DenoiseFilter`` for that.  ``NonLocalMeansParam`` (:122-157, the unfold / reflect-pad
variant with a learnable scalar ``h``, used nowhere in the reference) runs on a plain, untuned kernel
(``aisp_nlm_param_fwd``) for any odd window.

``rgb_to_luminance`` / ``ShiftStack`` / ``BoxFilter`` are the reference's small PyTorch helpers, kept for
API completeness (plain torch ops, not on any hot path).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from . import functional as AF
from ._lib import PSTRIDE

EPS = 1e-8


def rgb_to_luminance(rgb_tensor):
    """isp/denoise.py:11-17 (without the host-syncing range assert)."""
    rgb_tensor = torch.clip(rgb_tensor, 0.0, 1.0)
    return 0.299 * rgb_tensor[:, :1, ...] + 0.587 * rgb_tensor[:, 1:2, ...] + 0.114 * rgb_tensor[:, 2:, ...]


class ShiftStack(nn.Module):
    """isp/denoise.py:20-43."""

    def __init__(self, window_size):
        super().__init__()
        wx, wy = window_size if isinstance(window_size, (list, tuple)) else (window_size, window_size)
        assert wx % 2 == 1 and wy % 2 == 1, "window size must be odd"
        self.rx, self.ry = wx // 2, wy // 2

    def forward(self, tensor):
        out = []
        for x_shift in range(-self.rx, self.rx + 1):
            for y_shift in range(-self.ry, self.ry + 1):
                out.append(torch.roll(tensor, shifts=(y_shift, x_shift), dims=(2, 3)))
        return torch.stack(out, dim=-1)


class BoxFilter(nn.Module):
    """isp/denoise.py:46-65."""

    def __init__(self, window_size, reduction="mean"):
        super().__init__()
        wx, wy = window_size if isinstance(window_size, (list, tuple)) else (window_size, window_size)
        assert wx % 2 == 1 and wy % 2 == 1, "window size must be odd"
        self.rx, self.ry = wx // 2, wy // 2
        self.area = wx * wy
        self.reduction = reduction

    def forward(self, tensor):
        local_sum = torch.zeros_like(tensor)
        for x_shift in range(-self.rx, self.rx + 1):
            for y_shift in range(-self.ry, self.ry + 1):
                local_sum += torch.roll(tensor, shifts=(y_shift, x_shift), dims=(2, 3))
        return local_sum if self.reduction == "sum" else local_sum / self.area


class _NlmModule(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, h, gray: bool):
        _lib.require_image(rgb, "rgb")
        if ctx.needs_input_grad[0]:
            raise _lib.AispError("the bare NLM modules differentiate w.r.t. h only; use DenoiseFilter for d/d img")
        B, _, H, W = rgb.shape
        hb = h.reshape(-1).to(torch.float32)
        if hb.numel() not in (1, B):
            raise _lib.AispError(f"h must have 1 or {B} elements, got {tuple(h.shape)}")
        P = torch.zeros((B, PSTRIDE), dtype=torch.float32, device=rgb.device)
        P[:, 0] = hb if hb.numel() == B else hb.expand(B)
        ops = torch.full((B,), AF.OP_NLM, dtype=torch.int32, device=rgb.device)
        out = torch.empty_like(rgb)
        stash = torch.empty_like(rgb) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(rgb.device):
            rc = _lib.lib().aisp_nlm_module_fwd(rgb.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W,
                                                _lib.ptr(stash), int(gray), _lib.stream_ptr(rgb.device))
        _lib.check(rc, "aisp_nlm_module_fwd")
        ctx.save_for_backward(stash, ops)
        ctx.h_shape, ctx.dims = h.shape, (B, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        stash, ops = ctx.saved_tensors
        if stash is None:
            return None, None, None
        B, H, W = ctx.dims
        g = g.contiguous()
        gP = torch.zeros((B, PSTRIDE), dtype=torch.float32, device=g.device)
        sc = _lib.scratch(B, H, W, g.device)
        with torch.cuda.device(g.device):
            rc = _lib.lib().aisp_nlm_bwd(g.data_ptr(), stash.data_ptr(), ops.data_ptr(), B, H, W, gP.data_ptr(),
                                         sc.data_ptr(), sc.numel(), _lib.stream_ptr(g.device))
        _lib.check(rc, "aisp_nlm_bwd")
        gh = gP[:, 0]
        n = 1
        for d in ctx.h_shape:
            n *= d
        gh = gh.sum().reshape(ctx.h_shape) if n == 1 else gh.reshape(ctx.h_shape)
        return None, gh, None


class _NlmBase(nn.Module):
    GRAY = True

    def __init__(self, search_window_size=21, patch_size=7):
        super().__init__()
        if (search_window_size, patch_size) != (11, 5):
            raise _lib.AispError("the B200 NLM kernel is built for search_window_size=11, patch_size=5 "
                                 "(isp/filters.py:577), got (%s, %s)" % (search_window_size, patch_size))
        self.box_sum = BoxFilter(window_size=patch_size, reduction="sum")
        self.r = search_window_size // 2

    def forward(self, rgb, h):
        if not torch.is_tensor(h):
            h = torch.tensor([float(h)], dtype=torch.float32, device=rgb.device)
        return _NlmModule.apply(rgb, h, self.GRAY)


class NonLocalMeansGray(_NlmBase):
    """isp/denoise.py:93-119."""
    GRAY = True


class NonLocalMeans(_NlmBase):
    """isp/denoise.py:68-90 (per-channel distances and weights)."""
    GRAY = False


class _NlmParam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, h, window: int):
        _lib.require_image(rgb, "rgb")
        if ctx.needs_input_grad[0]:
            raise _lib.AispError("NonLocalMeansParam differentiates w.r.t. h only")
        B, _, H, W = rgb.shape
        luma = rgb_to_luminance(rgb).contiguous()
        out = torch.empty_like(rgb)
        stash = torch.empty_like(rgb) if ctx.needs_input_grad[1] else None
        hd = h.detach().reshape(-1)[:1].to(torch.float32).contiguous()
        with torch.cuda.device(rgb.device):
            rc = _lib.lib().aisp_nlm_param_fwd(rgb.data_ptr(), luma.data_ptr(), out.data_ptr(), _lib.ptr(stash),
                                               hd.data_ptr(), B, H, W, int(window), _lib.stream_ptr(rgb.device))
        _lib.check(rc, "aisp_nlm_param_fwd")
        ctx.save_for_backward(stash)
        ctx.h_shape = h.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (stash,) = ctx.saved_tensors
        if stash is None:
            return None, None, None
        return None, (g * stash).sum().reshape(ctx.h_shape), None


class NonLocalMeansParam(nn.Module):
    """isp/denoise.py:122-157: reflect-padded search window, patch box as large as the search window
    (:145-146), one learnable scalar ``h`` (``nn.Parameter`` of shape [1], as in the reference)."""

    def __init__(self, h0, search_window_size=21, patch_size=7):
        super().__init__()
        if search_window_size % 2 != 1:
            raise _lib.AispError("search_window_size must be odd")
        self.h = nn.Parameter(torch.tensor([float(h0)]), requires_grad=True)
        self.box_sum = BoxFilter(window_size=patch_size, reduction="sum")   # kept for parity; unused by forward (as in the reference)
        self.r = search_window_size // 2
        self.gen_window_stack = ShiftStack(window_size=search_window_size)
        self.search_window_size = search_window_size

    def forward(self, rgb):
        return _NlmParam.apply(rgb, self.h, self.search_window_size)
