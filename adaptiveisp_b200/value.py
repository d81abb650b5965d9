"""Drop-in critic (the reference's ``value.py``): same constructor, forward signature and ``state_dict``
keys.  It stays a PyTorch module -- conv/BN/FC nets are outside the hot path -- but its first step, the
``AdaptiveAvgPool2d`` of a full-resolution image (value.py:63), is taken from the block means that the
ISP kernels emitted from their store path when the image is an Agent output (``Agent.forward`` attaches
them as ``_aisp_down``), and from the one-pass block-mean kernel otherwise; the luminance / contrast /
saturation statistics (value.py:64-75) are computed from that 64x64 image exactly as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as AF
from .agent import FeatureExtractor as _AgentFeatureExtractor


class FeatureExtractor(nn.Module):
    """value.py:6-45: the Agent's conv pyramid without the dropout layer."""

    def __init__(self, shape=(17, 64, 64), mid_channels=32, output_dim=4096):
        super().__init__()
        self.output_dim = output_dim
        self.layers = _AgentFeatureExtractor(shape=shape, mid_channels=mid_channels, output_dim=output_dim,
                                             dropout_prob=0.0).layers

    def forward(self, x):
        return torch.reshape(self.layers(x), [-1, self.output_dim])


def value_statistics(images: torch.Tensor) -> torch.Tensor:
    """value.py:64-75 on the pooled image ``[B,3,h,w]`` -> ``[B,3]`` (luminance mean, luminance variance,
    mean saturation)."""
    lum = (images[:, 0, :, :] * 0.27 + images[:, 1, :, :] * 0.67 + images[:, 2, :, :] * 0.06 + 1e-5)[:, None, :, :]
    luminance = torch.mean(lum, dim=(1, 2, 3))
    contrast = torch.var(lum, dim=(1, 2, 3))
    clipped = torch.clip(images, min=0.0, max=1.0)
    i_max, _ = torch.max(clipped, dim=1)
    i_min, _ = torch.min(clipped, dim=1)
    sat = (i_max - i_min) / (torch.minimum(i_max + i_min, 2.0 - i_max - i_min) + 1e-2)
    saturation = torch.mean(sat, dim=[1, 2])
    return torch.stack([luminance, contrast, saturation], dim=1)


class Value(nn.Module):
    def __init__(self, cfg, shape=(19, 64, 64)):
        super().__init__()
        self.cfg = cfg
        self.feature_extractor = FeatureExtractor(shape=shape, mid_channels=cfg.base_channels,
                                                  output_dim=cfg.feature_extractor_dims)
        self.fc1 = nn.Linear(cfg.feature_extractor_dims, cfg.fc1_size)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self.fc2 = nn.Linear(cfg.fc1_size, 1)
        self.tanh = nn.Tanh()
        self.down_sample = nn.AdaptiveAvgPool2d((shape[1], shape[2]))

    def pooled(self, images):
        oh, ow = self.down_sample.output_size
        cached = getattr(images, "_aisp_down", None)
        if cached is not None and cached.shape == (images.shape[0], 3, oh, ow) and cached.device == images.device:
            return cached                      # emitted by the ISP kernels; differentiable w.r.t. the ISP step
        if images.is_cuda and not images.requires_grad and images.dtype == torch.float32 and images.is_contiguous() \
                and images.shape[2] % oh == 0 and images.shape[3] % ow == 0 and images.shape[0] * 3 <= 65535:
            return AF.block_mean(images, (oh, ow))
        return self.down_sample(images)

    def forward(self, images, states=None):
        images = self.pooled(images)
        if images.is_cuda and images.dtype == torch.float32 and images.is_contiguous() and \
                not (torch.is_grad_enabled() and images.requires_grad):
            state_feature = AF.value_stats(images)         # one launch; no gradient is needed through them
        else:
            state_feature = value_statistics(images)
        if states is None:
            states = state_feature
        else:
            assert len(states.shape) == len(state_feature.shape)
            states = torch.cat([states, state_feature], dim=1)
        states = states[:, :, None, None] + images[:, 0:1, :, :] * 0
        images = torch.cat([images, states], dim=1)
        feature = self.feature_extractor(images)
        return self.fc2(self.lrelu(self.fc1(feature)))
