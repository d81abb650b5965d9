"""Synthetic inputs shaped like what reaches the filters in the reference's LOD pipeline."""
import numpy as np
import torch


def lod_batch(B, H, W, seed=1234, letterbox=True, device="cpu"):
    """SURVEY.md §8(d) "LOD-shaped" synthetic frames: smooth low-frequency field x per-image
    brightness U(0.05,0.3) + shot/read noise (isp/unprocess_np.py:145-181 noise model), clipped,
    quantised to k/255, with exact-zero letterbox bars (3:2 content in a square frame)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand((B, 3, 8, 8), generator=g)
    field = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    bright = 0.05 + 0.25 * torch.rand((B, 1, 1, 1), generator=g)
    x = field * bright
    log_shot = np.log(1e-4) + (np.log(1.2e-2) - np.log(1e-4)) * torch.rand((B, 1, 1, 1), generator=g)
    shot = torch.exp(log_shot)
    log_read = 2.18 * log_shot + 1.2 + 0.26 * torch.randn((B, 1, 1, 1), generator=g)
    read = torch.exp(log_read)
    x = x + torch.sqrt(x * shot + read) * torch.randn((B, 3, H, W), generator=g)
    x = torch.round(torch.clip(x, 0.0, 1.0) * 255.0) / 255.0
    if letterbox:
        content = int(round(H * 2.0 / 3.0))
        top = (H - content) // 2
        x[:, :, :top, :] = 0.0
        x[:, :, top + content:, :] = 0.0
    return x.contiguous().to(device)
