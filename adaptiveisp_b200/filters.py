"""Drop-in filter classes: same names, constructor, attributes, methods, parameter layout and
``state_dict`` keys as the reference's ``isp/filters.py``, with the image arithmetic done by the
sm_100a kernels of ``libaisp_b200.so``.

Boundary: the reference looks filters up by class through ``config.cfg.filters``
(config.py:19-22), builds them as ``cls(cfg, predict=True).to(device)`` and registers them under
``get_short_name()`` (agent.py:72-75).  Swapping ``from isp.filters import *`` for
``from adaptiveisp_b200.filters import *`` in config.py is the whole integration.

What stays in PyTorch: the three tiny FC layers per filter and the feature->parameter regressors
(``[B,n]`` tensors; autograd chains the kernels' parameter gradients through them).
What runs in CUDA: ``process`` / ``forward`` / ``run`` / ``predict_param`` image math, forward and backward,
with the output clip of ``forward`` fused into the same pass.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as AF

__all__ = [
    "Filter", "ExposureFilter", "GammaFilter", "ImprovedWhiteBalanceFilter", "ColorFilter", "ToneFilter",
    "ToneFilterV2", "ContrastFilter", "WNBFilter", "SaturationPlusFilter", "DenoiseFilter", "SharpenUSMFilter",
    "SharpenFilter", "SharpenFilterV2", "CCMFilter", "FilterBank", "BankPredictor", "tanh01", "tanh_range", "lerp", "rgb2lum",
]


# -- small helpers kept for API compatibility (isp/filters.py:12-34) -----------------------------
def rgb2lum(image):
    lum = 0.27 * image[:, 0, :, :] + 0.67 * image[:, 1, :, :] + 0.06 * image[:, 2, :, :]
    return lum[:, None, :, :]


def lerp(a, b, l):
    return (1 - l) * a + l * b


def tanh01(x):
    return torch.tanh(x) * 0.5 + 0.5


def tanh_range(l, r, initial=None):
    shift = 0 if initial is None else math.atanh(2 * (initial - l) / (r - l) - 1)
    span = r - l

    def activation(x):
        return tanh01(x + shift) * span + l

    return activation


class Filter(nn.Module):
    """Base class; mirrors isp/filters.py:37-212.

    Subclasses set ``OP`` (the ``aisp_op`` code) and implement ``filter_param_regressor``.
    """

    OP: int = -1

    def __init__(self, cfg, short_name, num_filter_parameters, predict=False):
        super().__init__()
        self.cfg = cfg
        self.channels = 3
        self.num_filter_parameters = num_filter_parameters
        self.short_name = short_name
        self.filter_parameters = None
        if predict:
            self.fc1 = nn.Linear(cfg.feature_extractor_dims, cfg.fc1_size)
            self.lrelu = nn.LeakyReLU(negative_slope=0.2)
            self.fc_filter = nn.Linear(cfg.fc1_size, self.get_num_filter_parameters())
            self.fc_mask = nn.Linear(cfg.fc1_size, self.get_num_mask_parameters())
        self.predict = predict
        self._ones = {}

    # -- bookkeeping ---------------------------------------------------------------------------
    def get_short_name(self):
        assert self.short_name
        return self.short_name

    def get_num_filter_parameters(self):
        assert self.num_filter_parameters
        return self.num_filter_parameters

    def get_num_mask_parameters(self):
        return 6

    def use_masking(self):
        return False

    def debug_info_batched(self):
        return False

    def no_high_res(self):
        return False

    def extract_parameters(self, features):
        hidden = self.lrelu(self.fc1(features))
        return self.fc_filter(hidden), self.fc_mask(hidden)

    def filter_param_regressor(self, features):
        assert False

    # -- image math: CUDA ------------------------------------------------------------------------
    def _filter_apply(self, img, param, clip):
        return AF.apply_filter(img, param, self.OP, clip)

    def process(self, img, param):
        """Unclipped filter output (the stencil filters clip internally, as in the reference)."""
        return self._filter_apply(img, param, clip=False)

    def get_mask(self, img, mask_parameters=None):
        """Masking is disabled in the reference (``use_masking`` is False for every class,
        isp/filters.py:161-162, 170-173): the mask is the constant ones(1,1,1,1), cached per device
        here instead of being re-uploaded on every call."""
        if self.use_masking():
            raise NotImplementedError("spatial masking is dead code in the reference and is not part of the hot path")
        key = str(img.device)
        m = self._ones.get(key)
        if m is None:
            m = torch.ones((1, 1, 1, 1), dtype=torch.float32, device=img.device)
            self._ones[key] = m
        return m

    def _debug(self, filter_parameters):
        return filter_parameters if self.debug_info_batched() else filter_parameters[0]

    def forward(self, img, img_features=None, specified_parameter=None, high_res=None):
        """-> (low_res_output, high_res_output | None, debug_info); isp/filters.py:91-126.

        With the all-ones mask ``lerp(img, process(img,p), 1)`` is ``process(img,p)``; the final
        ``clip(.,0,1)`` is fused into the kernel (and its gradient mask into the backward kernel)."""
        if self.predict:
            assert (img_features is None) ^ (specified_parameter is None)
        if img_features is not None:
            filter_features, mask_parameters = self.extract_parameters(img_features)
            filter_parameters = self.filter_param_regressor(filter_features)
        else:
            assert not self.use_masking()
            filter_parameters = specified_parameter
            mask_parameters = torch.zeros(1, self.get_num_mask_parameters(), dtype=torch.float32)
        debug_info = {"filter_parameters": self._debug(filter_parameters)}
        # same values as the reference's side effect; detached because the mask is never used
        # (isp/filters.py:161-173) and a live reference would pin the whole autograd graph of this
        # call (and its AccumulateGrad nodes) between iterations, which breaks CUDA-graph capture
        self.mask_parameters = mask_parameters.detach()
        self.mask = self.get_mask(img, mask_parameters)
        debug_info["mask"] = self.mask[0]
        low_res_output = self._filter_apply(img, filter_parameters, clip=True)
        if high_res is not None:
            if self.no_high_res():
                high_res_output = high_res
            else:
                self.high_res_mask = self.get_mask(high_res, mask_parameters)
                high_res_output = self._filter_apply(high_res, filter_parameters, clip=True)
        else:
            high_res_output = None
        return low_res_output, high_res_output, debug_info

    def run(self, img, param):
        """isp/filters.py:128-139: no clip."""
        self.mask = self.get_mask(img)
        return self._filter_apply(img, param, clip=False)

    def run_v2(self, img, param):
        """isp/filters.py:141-152: ``param`` without the batch dim."""
        self.mask = self.get_mask(img)
        return self._filter_apply(img, param[None, :], clip=False)

    def predict_param(self, img, img_features):
        """isp/filters.py:154-159."""
        filter_features, _ = self.extract_parameters(img_features)
        filter_parameters = self.filter_param_regressor(filter_features)
        self.mask = self.get_mask(img)
        return self._filter_apply(img, filter_parameters, clip=False)

    # -- debug visualisation (host side, not on the hot path) ------------------------------------
    def _label(self, debug_info):
        p = debug_info["filter_parameters"].detach().float().cpu().numpy().reshape(-1)
        return self.get_short_name() + " " + " ".join("%+.2f" % v for v in p[:3])

    def visualize_filter(self, debug_info, canvas):
        import cv2
        text = self._label(debug_info)
        if canvas.shape[0] == 64:
            cv2.rectangle(canvas, (8, 40), (56, 52), (1, 1, 1), cv2.FILLED)
            cv2.putText(canvas, text, (8, 48), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (0, 0, 0))
        else:
            self.draw_high_res_text(text, canvas)

    def visualize_mask(self, debug_info, res):
        import cv2
        return cv2.resize(debug_info["mask"].cpu().numpy() * np.ones((1, 1, 3), dtype=np.float32),
                          dsize=res, interpolation=cv2.INTER_NEAREST)

    def draw_high_res_text(self, text, canvas):
        import cv2
        cv2.putText(canvas, text, (30, 128), cv2.FONT_HERSHEY_SIMPLEX, 0.8, (0, 0, 0), thickness=5)
        return canvas


class ExposureFilter(Filter):  # isp/filters.py:215-224
    OP = AF.OP_EXPOSURE

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "E", 1, predict)

    def filter_param_regressor(self, features):
        return tanh_range(-self.cfg.exposure_range, self.cfg.exposure_range, initial=0)(features)


class GammaFilter(Filter):  # isp/filters.py:235-245
    OP = AF.OP_GAMMA

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "G", 1, predict)

    def filter_param_regressor(self, features):
        bound = np.log(self.cfg.gamma_range)
        return torch.exp(tanh_range(-bound, bound)(features))


class ImprovedWhiteBalanceFilter(Filter):  # isp/filters.py:253-272
    OP = AF.OP_WB

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "W", 3, predict)
        self.num_filter_parameters = self.channels

    def filter_param_regressor(self, features):
        # the red feature is zeroed, so the red gain is 1 before the luminance normalisation
        # (built from device-side ops: the reference's numpy->tensor->.to(device) mask, isp/filters.py:260,
        # is a host->device copy on every call)
        masked = torch.cat([features[:, :1] * 0, features[:, 1:]], dim=1)
        gains = torch.exp(tanh_range(-0.5, 0.5)(masked))
        norm = 1.0 / (1e-5 + 0.27 * gains[:, 0] + 0.67 * gains[:, 1] + 0.06 * gains[:, 2])
        return gains * norm[:, None]


class ColorFilter(Filter):  # isp/filters.py:281-303
    OP = AF.OP_COLOR

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "C", 3 * cfg.curve_steps, predict)
        self.curve_steps = cfg.curve_steps
        assert cfg.curve_steps == 8, "the CUDA curve kernels are built for curve_steps == 8 (config.py:28)"

    def filter_param_regressor(self, features):
        curve = torch.reshape(features, shape=(-1, self.cfg.curve_steps, self.channels))[:, :, :, None, None]
        return tanh_range(*self.cfg.color_curve_range, initial=1)(curve)


class ToneFilter(Filter):  # isp/filters.py:326-347
    OP = AF.OP_TONE

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "T", cfg.curve_steps, predict)
        self.curve_steps = cfg.curve_steps
        assert cfg.curve_steps == 8, "the CUDA curve kernels are built for curve_steps == 8 (config.py:28)"

    def filter_param_regressor(self, features):
        curve = torch.reshape(features, shape=(-1, self.cfg.curve_steps, 1))[:, :, :, None, None]
        return tanh_range(*self.cfg.tone_curve_range)(curve)


class ToneFilterV2(ToneFilter):
    """isp/filters.py:365-387: ``process`` takes a flat ``[B,8]`` parameter.  (The reference's own
    predict path for this class is broken -- 5-D regressor output into a flat-param process -- and
    raises there; here both layouts are accepted because the kernel takes the flat row anyway.)"""


class ContrastFilter(Filter):  # isp/filters.py:406-419
    OP = AF.OP_CONTRAST

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "Ct", 1, predict)

    def filter_param_regressor(self, features):
        return torch.tanh(features)


class WNBFilter(Filter):  # isp/filters.py:427-437
    OP = AF.OP_WNB

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "BW", 1, predict)

    def filter_param_regressor(self, features):
        return torch.sigmoid(features)


class SaturationPlusFilter(Filter):  # isp/filters.py:536-560
    OP = AF.OP_SATPLUS

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "S+", 1, predict)

    def filter_param_regressor(self, features):
        return torch.sigmoid(features)


class DenoiseFilter(Filter):  # isp/filters.py:571-586  (NLM gray, 11x11 search, 5x5 patch)
    OP = AF.OP_NLM

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "NLM", 1, predict)

    def filter_param_regressor(self, features):
        return torch.sigmoid(features)


class SharpenUSMFilter(Filter):  # isp/filters.py:597-608
    OP = AF.OP_USM

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "USM", 2, predict)

    def filter_param_regressor(self, features):
        return tanh_range(*self.cfg.usm_sharpen_range)(features)


class SharpenFilter(Filter):  # isp/filters.py:621-631
    OP = AF.OP_SHARPEN

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "Shr", 1, predict)

    def filter_param_regressor(self, features):
        return tanh_range(*self.cfg.sharpen_range)(features)


class SharpenFilterV2(Filter):  # isp/filters.py:644-653
    OP = AF.OP_SHARPEN_V2

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "Shr", 1, predict)

    def filter_param_regressor(self, features):
        return tanh_range(*self.cfg.sharpen_range)(features)


class CCMFilter(Filter):  # isp/filters.py:694-708
    OP = AF.OP_CCM

    def __init__(self, cfg, predict=False):
        super().__init__(cfg, "CCM", 9, predict)

    def filter_param_regressor(self, features):
        return tanh_range(*self.cfg.ccm_range)(features)


class BankPredictor:
    """Features -> packed parameter rows ``[B,F,PSTRIDE]`` for a list of filter modules in a handful of
    launches: ONE GEMM for all ``fc1`` layers (their weights concatenated on the fly -- the modules keep
    their own parameters, so ``state_dict`` and the gradients are the per-module ones), the F small
    ``fc_filter`` GEMMs, and ONE kernel for every filter's ``filter_param_regressor``
    (``aisp_regress_fwd``; one more in backward).  The per-module statement (``extract_parameters`` +
    ``filter_param_regressor``: ~25 launches per filter and as many again in backward) stays available
    and is what this is tested against; ``fc_mask`` is not evaluated (its output is never used:
    isp/filters.py:161-173) and its parameters get no gradient, as in the reference."""

    def __init__(self, filters):
        self.filters = list(filters)
        self.ops = [int(f.OP) for f in self.filters]
        self.n = [int(f.get_num_filter_parameters()) for f in self.filters]
        self.offsets = [sum(self.n[:i]) for i in range(len(self.n))]
        self._dev = {}
        self._cfg_c = AF.regress_ranges(self.filters[0].cfg)

    def usable(self, features) -> bool:
        return features.is_cuda and features.dtype == torch.float32 and all(f.predict for f in self.filters)

    def _tables(self, device):
        t = self._dev.get(device)
        if t is None:
            t = (torch.tensor(self.ops, dtype=torch.int32, device=device),
                 torch.tensor(self.offsets, dtype=torch.int32, device=device))
            self._dev[device] = t
        return t

    def __call__(self, features):
        fl = self.filters
        F, B = len(fl), features.shape[0]
        w1 = torch.cat([f.fc1.weight for f in fl], dim=0)
        b1 = torch.cat([f.fc1.bias for f in fl], dim=0)
        hidden = nn.functional.leaky_relu(nn.functional.linear(features, w1, b1), 0.2).view(B, F, -1)
        raw = torch.cat([nn.functional.linear(hidden[:, i], f.fc_filter.weight, f.fc_filter.bias)
                         for i, f in enumerate(fl)], dim=1)
        fops, offs = self._tables(features.device)
        return AF.regress(raw, fops, offs, self._cfg_c, F)

    def split(self, packed):
        """Packed rows -> the per-filter parameter tensors in the reference's own layouts (views)."""
        out = []
        for i, f in enumerate(self.filters):
            p = packed[:, i, :self.n[i]]
            if f.OP == AF.OP_TONE:
                p = p.reshape(-1, 8, 1, 1, 1)
            elif f.OP == AF.OP_COLOR:
                p = p.reshape(-1, 8, 3, 1, 1)
            out.append(p)
        return out


class FilterBank:
    """Every filter of a list applied to the same batch in one banked launch set.

    Equivalent to the run-all loop of the reference's agent (agent.py:103-107)::

        stack = torch.stack([f(img, img_features)[0] for f in filters], dim=1)      # [B,F,3,H,W]

    but the image arithmetic of all F filters is three kernel launches (``functional.apply_bank``)
    instead of F, the image is fetched from HBM once, and the parameter prediction of all F filters is
    batched (:class:`BankPredictor`).  ``filters`` are the already constructed drop-in modules, so their
    weights / ``state_dict`` are untouched.  Not an ``nn.Module``: it owns no parameters.
    """

    def __init__(self, filters):
        self.filters = list(filters)
        if not self.filters:
            raise ValueError("empty filter bank")
        for f in self.filters:
            if f.use_masking():
                raise NotImplementedError("spatial masking is dead code in the reference and is not part of the hot path")
        self.ops = [int(f.OP) for f in self.filters]
        self.predictor = BankPredictor(self.filters)

    def parameters_for(self, img_features=None, specified_parameters=None):
        """-> list of per-filter parameter tensors in the reference's own layouts (the per-module
        statement: ``extract_parameters`` + ``filter_param_regressor`` of every filter)."""
        assert (img_features is None) ^ (specified_parameters is None)
        if specified_parameters is not None:
            assert len(specified_parameters) == len(self.filters)
            return list(specified_parameters)
        out = []
        for f in self.filters:
            feats, _ = f.extract_parameters(img_features)
            out.append(f.filter_param_regressor(feats))
        return out

    def packed_parameters(self, img_features=None, specified_parameters=None, batched=True):
        """-> (``P [B,F,PSTRIDE]``, per-filter parameter tensors)."""
        if img_features is not None and batched and self.predictor.usable(img_features):
            P = self.predictor(img_features)
            return P, self.predictor.split(P)
        params = self.parameters_for(img_features, specified_parameters)
        P = torch.stack([AF.pack_params(p, f.get_num_filter_parameters()) for p, f in zip(params, self.filters)], dim=1)
        return P, params

    def __call__(self, img, img_features=None, specified_parameters=None, clip=True, batched=True):
        """-> (stack ``[B,F,3,H,W]``, list of debug_info dicts as ``Filter.forward`` returns them)."""
        P, params = self.packed_parameters(img_features, specified_parameters, batched)
        stack = AF.apply_bank(img, P, self.ops, clip=clip)
        debug = []
        for p, f in zip(params, self.filters):
            debug.append({"filter_parameters": f._debug(p), "mask": f.get_mask(img)[0]})
        return stack, debug
