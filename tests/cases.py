"""Seeded edge-case inputs shared by the golden generator, the oracle tests and the GPU parity tests.

The vectors follow SURVEY.md Appendix B: exact-zero letterbox rows, saturated (==1.0) pixels,
values outside [0,1] (the ``run()`` path has no inter-stage clip), gray / two-channel-tie pixels for
the HSV code, pixels exactly on the tone-curve knots k/8, lattice (k/255) and off-lattice values.
"""
import numpy as np
import torch

from oracle import isp_oracle as O

ALL_OPS = [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_CCM, O.OP_SHARPEN, O.OP_NLM, O.OP_TONE, O.OP_CONTRAST,
           O.OP_SATPLUS, O.OP_WNB, O.OP_WB, O.OP_USM, O.OP_COLOR, O.OP_SHARPEN_V2]
POINTWISE_OPS = [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_CCM, O.OP_TONE, O.OP_CONTRAST, O.OP_SATPLUS, O.OP_WNB,
                 O.OP_WB, O.OP_COLOR]
STENCIL_OPS = [O.OP_SHARPEN, O.OP_SHARPEN_V2, O.OP_USM, O.OP_NLM]
AGENT_OPS = ALL_OPS[:10]  # config.py:19-22 order


def edge_image(B=3, H=20, W=24, seed=0, in_range=False):
    """[B,3,H,W] fp32.

    sample 0: k/255 lattice, low-light content, exact-zero letterbox rows top and bottom;
    sample 1: smooth floats in [0,1] with saturated patches and planted tie / knot pixels;
    sample 2+: floats spilling outside [0,1] (about [-0.15, 1.35]) unless ``in_range``.
    """
    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.rand((B, 3, H, W), generator=g)
    # sample 0: dark lattice + letterbox
    img[0] = torch.round(img[0] * 0.35 * 255.0) / 255.0
    bar = max(1, H // 6)
    img[0, :, :bar, :] = 0.0
    img[0, :, H - bar:, :] = 0.0
    if B > 1:
        s = img[1]
        s[:, 2:4, 2:5] = 1.0                      # saturated patch
        s[:, 5, 3] = 0.25                         # gray tie
        s[:, 5, 4] = 0.0                          # black pixel in content
        s[0, 6, 3], s[1, 6, 3], s[2, 6, 3] = 0.6, 0.6, 0.2    # R==G max
        s[0, 6, 4], s[1, 6, 4], s[2, 6, 4] = 0.1, 0.7, 0.7    # G==B max
        s[0, 6, 5], s[1, 6, 5], s[2, 6, 5] = 0.8, 0.3, 0.8    # R==B max
        s[0, 6, 6], s[1, 6, 6], s[2, 6, 6] = 0.3, 0.3, 0.9    # R==G min
        for k in range(9):                        # tone knots k/8 on every channel
            s[:, 7, k] = k / 8.0
        s[:, 8, 2] = 0.001                        # gamma clamp knee
        s[:, 8, 3] = 0.0005
    if B > 2 and not in_range:
        img[2:] = img[2:] * 1.5 - 0.15
    return img.contiguous()


def features(op, B, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(2000 + 17 * op + seed)
    return torch.randn((B, O.OP_NPARAMS[op]), generator=g) * scale


def params_for(op, B, seed=0):
    """Regressed parameters in the reference's own layout, from seeded N(0,1) features.

    CCM features are shrunk and biased towards a diagonally dominant matrix so that row sums stay
    away from 0 (the row-normalisation has no epsilon, isp/filters.py:707; the singular case is a
    separate test).
    """
    f = features(op, B, seed)
    if op == O.OP_CCM:
        f = 0.25 * f + torch.tensor([1.0, 0.1, -0.1, 0.05, 1.0, -0.05, -0.1, 0.1, 1.0])[None, :] * 0.6
    return f, O.regress(op, f)


def flat(op, p):
    """Reference-layout parameter -> flat [B,n] view."""
    return p.reshape(p.shape[0], O.OP_NPARAMS[op])


def grad_out(shape, seed=0):
    g = torch.Generator().manual_seed(3000 + seed)
    return torch.randn(shape, generator=g)


from adaptiveisp_b200.synthetic import lod_batch  # noqa: E402,F401  (re-exported for tests)
