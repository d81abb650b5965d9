"""CPU oracle for the AdaptiveISP differentiable filter chain.

TEST INFRASTRUCTURE ONLY.  Nothing in ``adaptiveisp_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may.  The product path is the CUDA library and fails loudly without it.

This is a restatement, in plain functional PyTorch (fp32, CPU), of the arithmetic of the
reference's ``process()`` / regressor / selection code.  Every function cites the reference
file:line it follows (paths relative to the reference checkout).  The ATen ops and their order
are kept the same as the reference's so that (i) CPU results are bit-comparable with the real
reference, (ii) ``torch.autograd`` over these functions yields the reference's gradients, and
(iii) timing this module on host cores is a faithful stand-in ("port") for the reference's
PyTorch CPU path on machines where the reference checkout is absent (the GPU box).

Parity pinning: the reference ships no tests / golden vectors for this path.
The oracle is therefore pinned against the *reference itself*, imported unmodified in the build
container by ``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this module against them.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# op codes: shared vocabulary with include/aisp_b200.h (first ten follow config.py:19-22 order)
# ----------------------------------------------------------------------------------------------
OP_EXPOSURE, OP_GAMMA, OP_CCM, OP_SHARPEN, OP_NLM, OP_TONE, OP_CONTRAST, OP_SATPLUS, OP_WNB, OP_WB, \
    OP_USM, OP_COLOR, OP_SHARPEN_V2 = range(13)

OP_NAMES = {
    OP_EXPOSURE: "E", OP_GAMMA: "G", OP_CCM: "CCM", OP_SHARPEN: "Shr", OP_NLM: "NLM", OP_TONE: "T",
    OP_CONTRAST: "Ct", OP_SATPLUS: "S+", OP_WNB: "BW", OP_WB: "W", OP_USM: "USM", OP_COLOR: "C",
    OP_SHARPEN_V2: "ShrV2",
}
OP_NPARAMS = {
    OP_EXPOSURE: 1, OP_GAMMA: 1, OP_CCM: 9, OP_SHARPEN: 1, OP_NLM: 1, OP_TONE: 8, OP_CONTRAST: 1,
    OP_SATPLUS: 1, OP_WNB: 1, OP_WB: 3, OP_USM: 2, OP_COLOR: 24, OP_SHARPEN_V2: 1,
}

CURVE_STEPS = 8  # config.py:28


# ----------------------------------------------------------------------------------------------
# helpers  (isp/filters.py:12-34)
# ----------------------------------------------------------------------------------------------
def lum_isp(img: torch.Tensor) -> torch.Tensor:
    """isp/filters.py:12-14 -- 0.27/0.67/0.06 luminance, keeps a singleton channel dim."""
    y = 0.27 * img[:, 0, :, :] + 0.67 * img[:, 1, :, :] + 0.06 * img[:, 2, :, :]
    return y[:, None, :, :]


def mix(a, b, t):
    """isp/filters.py:17-18."""
    return (1 - t) * a + t * b


def tanh01(x):
    """isp/filters.py:21-22."""
    return torch.tanh(x) * 0.5 + 0.5


def tanh_range(lo: float, hi: float, initial: Optional[float] = None):
    """isp/filters.py:25-34."""
    bias = 0 if initial is None else math.atanh(2 * (initial - lo) / (hi - lo) - 1)

    def act(x):
        return tanh01(x + bias) * (hi - lo) + lo

    return act


# ----------------------------------------------------------------------------------------------
# parameter regressors: raw FC features [B,n] -> filter parameters   (ranges: config.py:28-38)
# ----------------------------------------------------------------------------------------------
def regress(op: int, feat: torch.Tensor) -> torch.Tensor:
    """Feature -> parameter maps of every filter class.

    E   isp/filters.py:220-221      G  :240-242       W  :257-268      C  :287-291
    T   :332-335                    Ct :411-413       BW :432-433      S+ :542-543
    NLM :579-580                    USM:603-604       Shr:626-627      CCM:699-701
    Return shapes are the reference's (Tone is 5-D ``[B,8,1,1,1]``, Color ``[B,8,3,1,1]``).
    """
    if op == OP_EXPOSURE:
        return tanh_range(-3.5, 3.5, initial=0)(feat)
    if op == OP_GAMMA:
        lg = np.log(3)
        return torch.exp(tanh_range(-lg, lg)(feat))
    if op == OP_WB:
        keep = torch.tensor(np.array((0, 1, 1), dtype=np.float32).reshape(1, 3)).to(feat.device)
        s = torch.exp(tanh_range(-0.5, 0.5)(feat * keep))
        return s * (1.0 / (1e-5 + 0.27 * s[:, 0] + 0.67 * s[:, 1] + 0.06 * s[:, 2])[:, None])
    if op == OP_COLOR:
        c = torch.reshape(feat, shape=(-1, CURVE_STEPS, 3))[:, :, :, None, None]
        return tanh_range(0.90, 1.10, initial=1)(c)
    if op == OP_TONE:
        t = torch.reshape(feat, shape=(-1, CURVE_STEPS, 1))[:, :, :, None, None]
        return tanh_range(0.5, 2)(t)
    if op == OP_CONTRAST:
        return torch.tanh(feat)
    if op in (OP_WNB, OP_SATPLUS, OP_NLM):
        return torch.sigmoid(feat)
    if op == OP_USM:
        return tanh_range(0.0, 2.0)(feat)
    if op in (OP_SHARPEN, OP_SHARPEN_V2):
        return tanh_range(0.0, 10.0)(feat)
    if op == OP_CCM:
        return tanh_range(-2.0, 2.0)(feat)
    raise ValueError(f"unknown op {op}")


# ----------------------------------------------------------------------------------------------
# per-pixel filters: process(img, param)
# ----------------------------------------------------------------------------------------------
def exposure(img, p):
    """isp/filters.py:223-224.  p:[B,1]."""
    return img * torch.exp(p[:, :, None, None] * np.log(2))


def gamma(img, p):
    """isp/filters.py:244-245.  p:[B,1]."""
    return torch.pow(torch.clip(img, 0.001), p[:, :, None, None])


def white_balance(img, p):
    """isp/filters.py:270-271.  p:[B,3] (already luminance-normalised gains)."""
    return img * p[:, :, None, None]


def apply_ccm(img, m):
    """isp/filters.py:666-672.  img NCHW, m [B,3,3]; out_i = sum_j m[i,j] * img_j."""
    x = torch.permute(img, (0, 2, 3, 1))[:, :, :, None, :]
    out = torch.sum(x * m[:, None, None, :, :], dim=-1)
    return torch.permute(out, (0, 3, 1, 2))


def ccm(img, p):
    """isp/filters.py:703-708.  p:[B,9] row-major; rows normalised by their sum, no epsilon."""
    m = torch.reshape(p, shape=(-1, 3, 3))
    m = m / torch.sum(m, dim=-1, keepdim=True)
    return apply_ccm(img, m)


def _curve(img, knots, steps=CURVE_STEPS):
    """Shared body of ToneFilter.process (:337-347) and ColorFilter.process (:293-303).

    knots: [B,steps,1|3,1,1].  Piecewise-linear monotone curve, segments accumulated k=0..steps-1.
    """
    total = torch.sum(knots, dim=1) + 1e-30
    acc = img * 0
    for k in range(steps):
        acc += torch.clip(img - 1.0 * k / steps, 0, 1.0 / steps) * knots[:, k, :, :, :]
    acc *= steps / total
    return acc


def tone(img, p):
    """isp/filters.py:337-347.  p: [B,8,1,1,1] (or anything reshapeable to it)."""
    return _curve(img, p.reshape(p.shape[0], CURVE_STEPS, 1, 1, 1))


def color(img, p):
    """isp/filters.py:293-303.  p: [B,8,3,1,1], knot-major / channel-minor."""
    return _curve(img, p.reshape(p.shape[0], CURVE_STEPS, 3, 1, 1))


def contrast(img, p):
    """isp/filters.py:415-419.  p:[B,1]."""
    lum = torch.clip(lum_isp(img), 0.0, 1.0)
    target = -torch.cos(math.pi * lum) * 0.5 + 0.5
    boosted = img / (lum + 1e-6) * target
    return mix(img, boosted, p[:, :, None, None])


def wnb(img, p):
    """isp/filters.py:435-437.  p:[B,1]; desaturate towards (unclipped) luminance."""
    return mix(img, lum_isp(img), p[:, :, None, None])


def rgb_to_hsv(img):
    """isp/filters.py:445-478.  Ordered masked overwrites B, then G, then R (R wins ties)."""
    eps = 1e-8
    mx = img.max(1)[0]
    mn = img.min(1)[0]
    hue = torch.zeros((img.shape[0], img.shape[2], img.shape[3]), dtype=img.dtype, device=img.device)
    sel = img[:, 2] == mx
    hue[sel] = 4.0 + ((img[:, 0] - img[:, 1]) / (mx - mn + eps))[sel]
    sel = img[:, 1] == mx
    hue[sel] = 2.0 + ((img[:, 2] - img[:, 0]) / (mx - mn + eps))[sel]
    sel = img[:, 0] == mx
    hue[sel] = (0.0 + ((img[:, 1] - img[:, 2]) / (mx - mn + eps))[sel]) % 6
    hue[mn == mx] = 0.0
    hue = hue / 6
    sat = (mx - mn) / (mx + eps)
    sat[mx == 0] = 0
    return torch.cat([hue.unsqueeze(1), sat.unsqueeze(1), mx.unsqueeze(1)], dim=1)


def hsv_to_rgb(hsv):
    """isp/filters.py:481-533."""
    h, s, v = hsv[:, 0, :, :], hsv[:, 1, :, :], hsv[:, 2, :, :]
    h = h % 1
    s = torch.clamp(s, 0, 1)
    v = torch.clamp(v, 0, 1)
    r = torch.zeros_like(h)
    g = torch.zeros_like(h)
    b = torch.zeros_like(h)
    hi = torch.floor(h * 6)
    f = h * 6 - hi
    p = v * (1 - s)
    q = v * (1 - (f * s))
    t = v * (1 - ((1 - f) * s))
    table = ((v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q))
    for sextant, (rr, gg, bb) in enumerate(table):
        m = hi == sextant
        r[m] = rr[m]
        g[m] = gg[m]
        b[m] = bb[m]
    return torch.cat([r.unsqueeze(1), g.unsqueeze(1), b.unsqueeze(1)], dim=1)


def saturation_plus(img, p):
    """isp/filters.py:545-560.  p:[B,1]."""
    img = torch.clip(img, min=0.0, max=1.0)
    hsv = rgb_to_hsv(img)
    s = hsv[:, 1:2, :, :]
    v = hsv[:, 2:3, :, :]
    s2 = s + (1 - s) * (0.5 - torch.abs(0.5 - v)) * 0.8
    full = hsv_to_rgb(torch.cat([hsv[:, 0:1, :, :], s2, hsv[:, 2:, :, :]], dim=1))
    p = p[:, :, None, None]
    return img * (1.0 - p) + full * p


# ----------------------------------------------------------------------------------------------
# stencils
# ----------------------------------------------------------------------------------------------
def _blur3x3_keep_border(img):
    """isp/sharpen.py:119-138 (identical in :159-178): [[1,1,1],[1,5,1],[1,1,1]]/13 on the valid
    region, 1-px border keeps the input pixel."""
    ch = img.shape[1]
    k = torch.ones((3, 3), dtype=img.dtype, device=img.device)
    k[1, 1] = 5.0
    k /= k.sum()
    k = k.expand(img.shape[-3], 1, 3, 3)
    inner = F.conv2d(img, k, groups=ch)
    ones = torch.ones_like(inner)
    pad = [1, 1, 1, 1]
    return torch.where(F.pad(ones, pad) == 1, F.pad(inner, pad), img)


def sharpen(img, p):
    """SharpenFilter.process isp/filters.py:629-631 -> adjust_sharpness isp/sharpen.py:105-142."""
    f = p[:, :, None, None]
    out = img * f + _blur3x3_keep_border(img) * (1.0 - f)
    return torch.clip(out, 0.0, 1.0)


def sharpen_v2(img, p):
    """SharpenFilterV2.process isp/filters.py:651-653 -> sharpness isp/sharpen.py:145-182."""
    f = p[:, :, None, None]
    out = img + (img - _blur3x3_keep_border(img)) * f
    return torch.clip(out, 0.0, 1.0)


def _gauss1d(n: int, sigma: torch.Tensor):
    """isp/sharpen.py:15-23."""
    half = (n - 1) * 0.5
    x = torch.linspace(-half, half, steps=n).to(sigma.device)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def _gauss_blur5(img, sigma):
    """isp/sharpen.py:63-81 with kernel_size=(5,5): reflect pad 2, depthwise conv."""
    k1 = _gauss1d(5, sigma).to(img.device, dtype=img.dtype)
    k2 = torch.mm(k1[:, None], k1[None, :])
    k2 = k2.expand(img.shape[-3], 1, 5, 5)
    squeeze = img.ndim < 4
    x = img.unsqueeze(0) if squeeze else img
    x = F.pad(x, [2, 2, 2, 2], mode="reflect")
    x = F.conv2d(x, k2, groups=x.shape[-3])
    return x.squeeze(0) if squeeze else x


def usm(img, p):
    """SharpenUSMFilter.process isp/filters.py:606-608 -> unsharp_mask isp/sharpen.py:84-102.

    p:[B,2] = (sigma, amount).  B>1 takes the reference's per-image loop, B==1 its batched branch.
    """
    sigma, amount = p[:, 0], p[:, 1]
    if img.ndim > 3 and sigma.shape[0] > 1:
        out = torch.zeros_like(img)
        for b in range(img.shape[0]):
            blurred = _gauss_blur5(img[b], sigma[b])
            out[b] = img[b] + (img[b] - blurred) * amount[b]
    else:
        blurred = _gauss_blur5(img, sigma.squeeze(-1))
        out = img + (img - blurred) * amount
    return torch.clip(out, 0.0, 1.0)


def lum_nlm(rgb):
    """isp/denoise.py:11-17 (0.299/0.587/0.114 on the clipped image)."""
    rgb = torch.clip(rgb, 0.0, 1.0)
    return 0.299 * rgb[:, :1, ...] + 0.587 * rgb[:, 1:2, ...] + 0.114 * rgb[:, 2:, ...]


def _box_sum(t, radius):
    """isp/denoise.py:46-65 with reduction='sum' (circular: torch.roll)."""
    acc = torch.zeros_like(t)
    for xs in range(-radius, radius + 1):
        for ys in range(-radius, radius + 1):
            acc += torch.roll(t, shifts=(ys, xs), dims=(2, 3))
    return acc


def nlm_gray(img, p, search: int = 11, patch: int = 5):
    """DenoiseFilter.process isp/filters.py:582-586 -> NonLocalMeansGray isp/denoise.py:93-119.

    p:[B,1] = h.  All boundaries circular; x-shift outer loop, y-shift inner.
    """
    rgb = torch.clip(img, min=0.0, max=1.0)
    h = p[:, :, None, None]
    r = search // 2
    wsum = torch.zeros((rgb.shape[0], 1, rgb.shape[2], rgb.shape[3])).float().to(rgb.device)
    acc = torch.zeros_like(rgb)
    y = lum_nlm(rgb)
    for xs in range(-r, r + 1):
        for ys in range(-r, r + 1):
            rgb_s = torch.roll(rgb, shifts=(ys, xs), dims=(2, 3))
            y_s = torch.roll(y, shifts=(ys, xs), dims=(2, 3))
            dist = torch.sqrt(torch.relu(_box_sum((y - y_s) ** 2, patch // 2)))
            w = torch.exp(-dist / (torch.relu(h) + 1e-8))
            acc += rgb_s * w
            wsum += w
    return torch.clamp(acc / wsum, 0.0, 1.0)


def nlm_gray_module(rgb, h, search: int = 11, patch: int = 5):
    """The bare ``NonLocalMeansGray(search, patch).forward(rgb, h)`` of isp/denoise.py:93-119 (no filter
    wrapper): distances on the luma of the CLIPPED image (``rgb_to_luminance`` clips, :14), averages of
    the image as given.  h broadcastable to ``[B,1,1,1]``."""
    r = search // 2
    wsum = torch.zeros((rgb.shape[0], 1, rgb.shape[2], rgb.shape[3]), dtype=rgb.dtype).to(rgb.device)
    acc = torch.zeros_like(rgb)
    y = lum_nlm(torch.clip(rgb, 0.0, 1.0))
    for xs in range(-r, r + 1):
        for ys in range(-r, r + 1):
            rgb_s = torch.roll(rgb, shifts=(ys, xs), dims=(2, 3))
            y_s = torch.roll(y, shifts=(ys, xs), dims=(2, 3))
            dist = torch.sqrt(torch.relu(_box_sum((y - y_s) ** 2, patch // 2)))
            w = torch.exp(-dist / (torch.relu(h) + 1e-8))
            acc += rgb_s * w
            wsum += w
    return torch.clamp(acc / wsum, 0.0, 1.0)


def nlm_rgb_module(rgb, h, search: int = 11, patch: int = 5):
    """The bare ``NonLocalMeans(search, patch).forward(rgb, h)`` of isp/denoise.py:68-90: per-channel
    patch distances and weights ``[B,3,H,W]`` on the image as given (no clip of the input)."""
    r = search // 2
    wsum = torch.zeros_like(rgb)
    acc = torch.zeros_like(rgb)
    for xs in range(-r, r + 1):
        for ys in range(-r, r + 1):
            rgb_s = torch.roll(rgb, shifts=(ys, xs), dims=(2, 3))
            dist = torch.sqrt(torch.relu(_box_sum((rgb - rgb_s) ** 2, patch // 2)))
            w = torch.exp(-dist / (torch.relu(h) + 1e-8))
            acc += rgb_s * w
            wsum += w
    return torch.clamp(acc / wsum, 0.0, 1.0)


def nlm_param_module(rgb, h, search: int = 21):
    """``NonLocalMeansParam(h0, search).forward(rgb)`` of isp/denoise.py:122-157, shift by shift instead of
    through the unfolded window stacks: the luma and the image are padded by reflection (:133-137), the
    squared luma differences are formed at the unpadded positions (:142), padded by reflection again
    (:144) and box-summed over a window as large as the SEARCH window (:145-146); one scalar ``h``."""
    import torch.nn.functional as F
    r = (search - 1) // 2
    B, _, H, W = rgb.shape
    y = lum_nlm(torch.clip(rgb, 0.0, 1.0))
    pad = [r, r, r, r]
    y_pad = F.pad(y, pad=pad, mode="reflect")
    rgb_pad = F.pad(rgb, pad=pad, mode="reflect")
    hh = torch.relu(h) + 1e-8
    acc = torch.zeros_like(rgb)
    wsum = torch.zeros_like(y)
    for wy in range(search):            # window index = wy * search + wx (unfold order)
        for wx in range(search):
            dis = (y - y_pad[:, :, wy:wy + H, wx:wx + W]) ** 2
            box = F.avg_pool2d(F.pad(dis, pad=pad, mode="reflect"), search, stride=1) * float(search * search)
            w = torch.exp(-torch.sqrt(torch.relu(box)) / hh)
            acc = acc + w * rgb_pad[:, :, wy:wy + H, wx:wx + W]
            wsum = wsum + w
    return torch.clamp(acc / wsum, 0.0, 1.0)


def shot_read_noise(image, shot, read, z, gain=None):
    """isp/unprocess_np.py:131-138,178-181 with the standard normals ``z`` injected: brightness ratio, then
    ``image + sqrt(image * shot + read) * z`` (what ``np.random.normal(0, sqrt(variance))`` computes from the
    normals it draws).  numpy arrays, float64 like the reference; per-image levels broadcast from ``[B]``."""
    import numpy as np
    image = np.asarray(image, dtype=np.float64)
    bshape = (image.shape[0],) + (1,) * (image.ndim - 1)
    if gain is not None:
        image = image * np.asarray(gain, dtype=np.float64).reshape(bshape)
    variance = image * np.asarray(shot, dtype=np.float64).reshape(bshape) + np.asarray(read, dtype=np.float64).reshape(bshape)
    return image + np.sqrt(variance) * np.asarray(z, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# dispatch + Filter.forward / Filter.run wrappers
# ----------------------------------------------------------------------------------------------
_PROCESS = {
    OP_EXPOSURE: exposure, OP_GAMMA: gamma, OP_CCM: ccm, OP_SHARPEN: sharpen, OP_NLM: nlm_gray,
    OP_TONE: tone, OP_CONTRAST: contrast, OP_SATPLUS: saturation_plus, OP_WNB: wnb, OP_WB: white_balance,
    OP_USM: usm, OP_COLOR: color, OP_SHARPEN_V2: sharpen_v2,
}


def process(op: int, img: torch.Tensor, param: torch.Tensor) -> torch.Tensor:
    """``Filter.process`` of the class with op code ``op``; ``param`` as the regressor returns it
    (or flat ``[B,n]``)."""
    return _PROCESS[op](img, param)


def run(op: int, img, param):
    """Filter.run isp/filters.py:128-139: lerp(img, process, ones) and NO clip."""
    one = torch.ones((1, 1, 1, 1), dtype=torch.float32).to(img.device)
    return mix(img, process(op, img, param), one)


def forward(op: int, img, param):
    """Filter.forward isp/filters.py:91-126 with masking disabled: clip(lerp(img, process, 1), 0, 1)."""
    return torch.clip(run(op, img, param), 0.0, 1.0)


def chain(ops: Sequence[int], img, params: Sequence[torch.Tensor], clip_each: bool):
    """A fixed filter sequence (isp/filters.py:753-815 uses run(); Agent rollouts use forward())."""
    x = img
    for op, p in zip(ops, params):
        x = forward(op, x, p) if clip_each else run(op, x, p)
    return x


# ----------------------------------------------------------------------------------------------
# Agent-side selection / gather / state update   (agent.py)
# ----------------------------------------------------------------------------------------------
STATE_REWARD_DIM, STATE_STOPPED_DIM, STATE_STEP_DIM, STATE_DROPOUT_BEGIN = 0, 1, 2, 3  # util.py:15-18


def pdf_sample(pdf, u):
    """agent.py:12-16."""
    pdf = pdf / (torch.sum(pdf, dim=1, keepdim=True) + 1e-36)
    cdf = torch.cumsum(pdf, dim=1) - pdf
    return torch.sum(torch.less(cdf, u).to(torch.int32), dim=1) - 1


def one_hot(n, index):
    """agent.py:18-23 -> int64 [B,n]."""
    lab = torch.zeros((n, *index.shape), dtype=torch.int64, device=index.device)
    for i in range(n):
        lab[i, index == i] = 1
    return lab.permute(1, 0)


def mix_pdf(logits, exploration=0.05):
    """agent.py:126-132: softmax + 1e-37, exploration mix, renormalise."""
    n = logits.shape[1]
    pdf = torch.softmax(logits, dim=1) + 1e-37
    pdf = pdf * (1 - exploration) + exploration * 1.0 / n
    return pdf / (torch.sum(pdf, dim=1, keepdim=True) + 1e-30)


def select(pdf, u, training: bool, forced: Optional[int] = None):
    """agent.py:138-143 -> int64 [B]."""
    rnd = pdf_sample(pdf, u)
    mx = torch.argmax(pdf, dim=1).to(torch.int32)
    if forced is not None:
        return torch.from_numpy(np.array([forced] * mx.shape[0])).to(torch.int64)
    t = 1 if training else 0
    return (t * rnd + (1 - t) * mx).to(torch.int64)


def gather_selected(stack, hot):
    """agent.py:154: sum_f stack[:,f] * one_hot[:,f]."""
    return torch.sum(stack * hot[:, :, None, None, None], dim=1)


def next_states(states, hot, test_steps=5):
    """agent.py:234-259 -> (new_states, usage_penalty, early_stop_penalty(=0 by construction))."""
    last = (torch.abs(states[:, STATE_STEP_DIM:STATE_STEP_DIM + 1] + 1 - test_steps) < 1e-4).to(torch.float32)
    step = (states[:, STATE_STEP_DIM] + 1)[:, None]
    usage = states[:, STATE_STEP_DIM + 1:]
    usage_pen = torch.sum(usage * hot, dim=1, keepdim=True)
    new_usage = torch.maximum(usage, hot)
    return torch.cat([last, last, step, new_usage], dim=1), usage_pen
