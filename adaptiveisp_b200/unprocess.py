"""Noise model of the reference's synthetic-RAW pipeline (``isp/unprocess_np.py:131-181``) on the GPU.

The reference degrades sRGB frames on the host with NumPy before they enter the hot path: a random
brightness scale (``adjust_random_brightness``), per-image noise levels from a log-log linear camera model
(``random_noise_levels_log`` / ``_linear``) and heteroscedastic Gaussian noise (``add_read_and_shot_noise``).
Here the per-image scalars are still drawn on the host (a handful of numbers, same formulas and the same
NumPy RNG calls, so a seeded ``np.random`` reproduces the reference's levels), the per-element work is ONE
streaming kernel over the resident batch (``aisp_shot_read_noise``: 8 B per element with in-kernel Philox
normals, 12 B when the caller supplies them).  The rest of ``unprocess_np.py`` (inverse tone / gamma / CCM /
gains, mosaic) is host-side data synthesis outside the path and is not built.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def adjust_random_brightness_ratio(s_range=(0.1, 0.3)) -> float:
    """The ratio ``adjust_random_brightness`` (isp/unprocess_np.py:131-138) multiplies the image with."""
    if isinstance(s_range, (list, tuple)):
        assert s_range[0] < s_range[1], "s_range[0] should less than s_range[1]"
        return float(np.random.rand() * (s_range[1] - s_range[0]) + s_range[0])
    return float(s_range)


def random_noise_levels_log(shot_noise=None):
    """isp/unprocess_np.py:146-159: shot noise log-uniform in [1e-4, 1.2e-2], read noise on the line
    ``log(read) = 2.18 log(shot) + 1.20`` plus N(0, 0.26)."""
    if shot_noise is None:
        log_shot_noise = np.random.uniform(np.log(0.0001), np.log(0.012))
        shot_noise = np.exp(log_shot_noise)
    else:
        log_shot_noise = np.log(shot_noise)
    log_read_noise = 2.18 * log_shot_noise + 1.20 + np.random.normal(0, 0.26)
    return shot_noise, np.exp(log_read_noise)


def random_noise_levels_linear(shot_noise=None):
    """isp/unprocess_np.py:162-175: as above with the shot noise uniform in [1e-4, 1.2e-2]."""
    if shot_noise is None:
        shot_noise = np.random.uniform(0.0001, 0.012)
    log_shot_noise = np.log(shot_noise)
    log_read_noise = 2.18 * log_shot_noise + 1.20 + np.random.normal(0, 0.26)
    return shot_noise, np.exp(log_read_noise)


def _per_image(v, B, device):
    t = torch.as_tensor(v, dtype=torch.float32).reshape(-1)
    if t.numel() == 1:
        t = t.expand(B)
    if t.numel() != B:
        raise _lib.AispError(f"expected 1 or {B} per-image values, got {t.numel()}")
    return t.contiguous().to(device)


@torch.no_grad()
def add_read_and_shot_noise(image: torch.Tensor, shot_noise=0.01, read_noise=0.005, *, gain=None, z=None,
                            seed: int = 0, offset: int = 0, out=None) -> torch.Tensor:
    """``image + N(0, sqrt(image * shot + read))`` (isp/unprocess_np.py:178-181) for a resident batch
    ``[B, ...]``; ``shot_noise`` / ``read_noise`` / ``gain`` scalars or one value per image (``gain``: the
    brightness ratio applied first).  ``z``: standard normals of the image's shape to use instead of the
    in-kernel Philox stream (``seed``, ``offset``: a call consumes ``ceil(n_per_image / 4)`` counters)."""
    if not image.is_cuda or image.dtype != torch.float32 or not image.is_contiguous():
        raise _lib.AispError("image must be a contiguous CUDA float32 tensor")
    B = image.shape[0]
    n = image.numel() // B
    dev = image.device
    sh, rd = _per_image(shot_noise, B, dev), _per_image(read_noise, B, dev)
    gn = None if gain is None else _per_image(gain, B, dev)
    if z is not None and (z.shape != image.shape or not z.is_cuda or z.dtype != torch.float32 or not z.is_contiguous()):
        raise _lib.AispError("z must be a contiguous CUDA float32 tensor of the image's shape")
    out = torch.empty_like(image) if out is None else out
    with torch.cuda.device(dev):
        rc = _lib.lib().aisp_shot_read_noise(image.data_ptr(), _lib.ptr(z), out.data_ptr(), sh.data_ptr(), rd.data_ptr(),
                                             _lib.ptr(gn), B, n, int(seed), int(offset), _lib.stream_ptr(dev))
    _lib.check(rc, "aisp_shot_read_noise")
    return out
